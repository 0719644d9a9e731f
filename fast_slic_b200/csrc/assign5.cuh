// fast_slic_b200/csrc/assign5.cuh -- the assign + update hot loop, TMA-staged (round 2).
//
// Same contract as k_assign_warp (assign.cuh; /root/reference/src/context.cpp:200-298 assign / assign_clusters,
// :302-387 update; AVX2 form arch/x64/avx2.h:11-185): every pixel of the active rows takes the lexicographic
// minimum of (d, phase, k) over the clusters whose (2S+1)^2 window covers it, pixels no window covers keep
// their label, and the per-cluster sums of the update are accumulated exactly.  What changed is where the
// instructions go.  k_assign_warp spent 77 % of its issue slots outside the distance loop (ncu, round 1):
// per-tile candidate lists built by 8 lanes each, 64-bit address arithmetic and bounds predicates around
// every LDG/STG, a four-way one-hot compare per MMA.  Here:
//
//   * quad tiles arrive by TMA: one elected lane per warp issues cp.async.bulk.tensor.3d (3-D tensor map
//     (x, sub-row, image) whose row stride is `stride` image rows, so the sub-sampled rows land as a dense
//     [4][32*TPS] tile in the warp's shared block); completion on a per-warp mbarrier; the load of the
//     NEXT super tile is in flight while this one is processed.  Out-of-image parts of a box are zero-filled
//     by the hardware: no bounds predicates on the loads.
//   * labels leave by TMA: lanes put u16 labels into a [4][32*TPS] staging tile, one bulk tensor store per
//     super tile writes them (clipped by the hardware at the image edge).  Super tiles that contain a pixel
//     no window covers -- or that the image edge clips -- take a per-pixel store path instead, which also
//     implements "keep the previous label" (context.cpp:289-294 never fires for such pixels).
//   * ONE candidate list per super tile (32*TPS columns x 4 sub-rows): <= 32 candidates, one lane each,
//     ranked once by (phase, k) with vector loads of the keys; the per-tile lists are derived from it with a
//     REDUX.OR bit mask per tile (bit = rank), so a candidate's position in a tile list is a popcount and
//     the lists stay in visiting order without any per-tile sort.
//   * update: the one-hot operand needs 3 instead of 4 logic ops per word (rank bytes stay below 0x80),
//     and the second half of the MMA A operand (candidates 8..15) is skipped when a tile has <= 8.
//
// HBM per processed pixel: 4 B quad read + 2 B label written, as before.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

#include "assign.cuh"

#define A5_MAXWARPS 32
#define A5_WBLK 5376     // bytes of shared memory per warp (multiple of 128: TMA destinations need 128-byte alignment)
#define A5_OFF_QUAD 0    // [4][32*TPS] u32  quad tile                      (<= 2048 B)
#define A5_OFF_LAB 2048  // [4][32*TPS] u16  label staging tile             (<= 1024 B)
#define A5_OFF_ENT 3072  // [4][32] uint2    per tile {colour, patch offset} in visiting order (1024 B)
#define A5_OFF_TK 4096   // [4][32] u16      per tile cluster number in visiting order         (256 B)
#define A5_OFF_SCR 4352  // 1024 B scratch: CInfo[32] + keys[32] while the list is built; MMA staging [32][8] u32 afterwards
#define A5_BIGKEY (FSLIC_BIGSP << 16)

__device__ __forceinline__ uint32_t a5_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void a5_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void a5_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a5_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "A5_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra A5_DONE;\n"
        "bra A5_WAIT;\n"
        "A5_DONE:\n"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void a5_tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void a5_tma_store_3d(const CUtensorMap* map, uint32_t src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(x), "r"(y), "r"(z)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void a5_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void a5_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per-byte equality of two words whose bytes are all below 0x80 -> 0x80 in every equal byte (3 logic ops)
__device__ __forceinline__ uint32_t eq7(uint32_t w, uint32_t m) { return ~((w ^ m) + 0x7f7f7f7fu) & 0x80808080u; }

template <int TS, int STRIDE, bool UPDATE, int TPS>
__global__ void __launch_bounds__(32 * A5_MAXWARPS, 1)
    k_assign5(AssignParams ap, const __grid_constant__ CUtensorMap tm_quad, const __grid_constant__ CUtensorMap tm_lab,
              const uint32_t* __restrict__ quad, uint16_t* __restrict__ labels, const CInfo* __restrict__ cinfo,
              const int* __restrict__ cell_start, unsigned long long* __restrict__ acc,
              const uint16_t* __restrict__ g_tbl) {
    constexpr int R = 4;                      // sub-rows per tile (one per register of a lane)
    constexpr int BW = 32 * TPS;              // columns of a super tile == TMA box width
    constexpr uint32_t QBYTES = BW * R * 4;   // bytes one quad box delivers (out-of-image parts included: zero fill)
    static_assert(STRIDE >= 1, "the sub-row pitch is an immediate of the patch loads");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const uint32_t tbl_bytes = ((uint32_t)ap.tbl_elems * 2u + 127u) & ~127u;
    unsigned char* wblk = smem_raw + tbl_bytes + warp * A5_WBLK;
    const uint32_t* s_quad = reinterpret_cast<const uint32_t*>(wblk + A5_OFF_QUAD);
    uint16_t* s_lab = reinterpret_cast<uint16_t*>(wblk + A5_OFF_LAB);
    uint2 (*s_ent)[32] = reinterpret_cast<uint2 (*)[32]>(wblk + A5_OFF_ENT);
    uint16_t (*s_tk)[32] = reinterpret_cast<uint16_t (*)[32]>(wblk + A5_OFF_TK);
    CInfo* s_u = reinterpret_cast<CInfo*>(wblk + A5_OFF_SCR);
    uint32_t* s_key = reinterpret_cast<uint32_t*>(wblk + A5_OFF_SCR + 512);
    uint32_t (*s_feat)[8] = reinterpret_cast<uint32_t (*)[8]>(wblk + A5_OFF_SCR);
    const uint32_t bar = a5_smem(smem_raw + tbl_bytes + nwarps * A5_WBLK + warp * 8);
    const uint32_t quad_dst = a5_smem(wblk + A5_OFF_QUAD), lab_src = a5_smem(wblk + A5_OFF_LAB);

    if (lane == 0) a5_mbar_init(bar, 1);
    {   // the spatial patch, once per CTA (16-byte copies; the table is padded to a multiple of 8 elements)
        const uint4* src = reinterpret_cast<const uint4*>(g_tbl);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw);
        for (int t = tid; t < (ap.tbl_elems + 7) / 8; t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();

    const int S = ap.S, W = ap.W, H = ap.H;
    constexpr int stride = STRIDE;
    const int g = lane >> 2, tig = lane & 3;  // MMA fragment coordinates
    const uint32_t lt_mask = (1u << lane) - 1u;

    // super-tile walk without divisions: (b, ty, sx) advance by a fixed (db, dty, dsx) with carries
    const int stx = (ap.tiles_x + TPS - 1) / TPS;  // super tiles per tile row
    const long per_img = (long)stx * ap.tiles_y;
    const long total = per_img * ap.B;
    const long wstride = (long)gridDim.x * nwarps;
    const long first = (long)blockIdx.x * nwarps + warp;
    int b = (int)(first / per_img);
    const int tl0 = (int)(first - (long)b * per_img);
    int ty = tl0 / stx, sx = tl0 - ty * stx;
    const int db = (int)(wstride / per_img);
    const int dtl = (int)(wstride - (long)db * per_img);
    const int dty = dtl / stx, dsx = dtl - dty * stx;

    uint32_t phase = 0;
    bool store_pending = false;  // warp uniform: a bulk store of this warp's staging tile may still be reading it
    if (first < total && lane == 0) {
        a5_mbar_expect_tx(bar, QBYTES);
        a5_tma_load_3d(quad_dst, &tm_quad, sx * BW, ty * R, b, bar);
    }

    for (long st = first; st < total; st += wstride) {
        // the super tile after this one (its quad box is requested while this one is processed)
        int nb = b + db, nty = ty + dty, nsx = sx + dsx;
        if (nsx >= stx) { nsx -= stx; nty += 1; }
        if (nty >= ap.tiles_y) { nty -= ap.tiles_y; nb += 1; }
        const bool has_next = st + wstride < total;

        const int wsr0 = ty * R;
        const int nrow = min(R, ap.nsub - wsr0);  // valid sub-rows of this tile row (>= 1)
        const int wi0 = ap.rem + wsr0 * stride, wi1 = wi0 + (nrow - 1) * stride;
        const int sj0 = sx * BW, sj1 = min(sj0 + BW, W) - 1;
        const int ntile = min(TPS, ap.tiles_x - sx * TPS);  // tiles of this super tile that start inside the image
        const bool edge = (sj0 + BW > W) || (nrow < R);
        const CInfo* ci = cinfo + (size_t)b * ap.K;
        const int* cs = cell_start + (size_t)b * (ap.ncell + 1);
        unsigned long long* ac = acc + (size_t)b * ap.K * 4;

        // ---- L. the candidate list of the super tile: clusters with cy in [wi0-S, wi1+S], cx in [sj0-S, sj1+S] ----
        int n = 0;            // candidates found (warp uniform)
        uint32_t ncnt = 0;    // per-tile list lengths, one byte each
        {
            const int ylo = wi0 - S, yhi = wi1 + S, xlo = sj0 - S, xhi = sj1 + S;
            const int cr0 = div_g(max(ylo, 0), ap.Ginv), cr1 = div_g(min(yhi, H - 1), ap.Ginv);
            const int cc0 = div_g(max(xlo, 0), ap.Ginv), cc1 = div_g(min(xhi, W - 1), ap.Ginv);
            const int nr = cr1 - cr0 + 1;  // cell rows: one contiguous range of the cell-sorted records each
            int rs = 0, cnt = 0;
            if (lane < nr) {
                const int* p = cs + (cr0 + lane) * ap.cellW;
                rs = p[cc0];
                cnt = p[cc1 + 1] - rs;
            }
            int incl = cnt;  // inclusive scan over the nr ranges
            for (int o = 1; o < nr; o <<= 1) {
                const int y = __shfl_up_sync(FSLIC_FULL, incl, o);
                if (lane >= o) incl += y;
            }
            const int T = __shfl_sync(FSLIC_FULL, incl, min(nr, 32) - 1);
            if (nr > 32) n = 33;  // cannot happen for G >= S (at most ~9 cell rows); treated like an overflow
            for (int t0 = 0; t0 < T && n <= 32; t0 += 32) {
                const int t = t0 + lane;
                int row = 0;  // range of flat index t: the first one whose inclusive prefix exceeds t
                for (int r = 0; r < nr - 1; r++) row += (t >= __shfl_sync(FSLIC_FULL, incl, r));
                const int rincl = __shfl_sync(FSLIC_FULL, incl, row);
                const int rcnt = __shfl_sync(FSLIC_FULL, cnt, row);
                const int rstart = __shfl_sync(FSLIC_FULL, rs, row);
                bool hit = false;
                CInfo rec;
                if (t < T) {
                    rec = ci[rstart + (t - (rincl - rcnt))];
                    const int cy = (int16_t)(rec.cyx & 0xffff), cx = rec.cyx >> 16;
                    hit = (cy >= ylo) && (cy <= yhi) && (cx >= xlo) && (cx <= xhi);
                }
                const unsigned m = __ballot_sync(FSLIC_FULL, hit);
                const int slot = n + __popc(m & lt_mask);
                if (hit && slot < 32) {
                    s_u[slot] = rec;
                    s_key[slot] = rec.sortkey;
                }
                n += __popc(m);
            }
            // pad the key vector to a multiple of 4 with +infinity (keys are phase << 16 | k < 2^18, all distinct)
            if (n <= 32 && lane >= n && lane < ((n + 3) & ~3)) s_key[lane] = 0xffffffffu;
            __syncwarp();
            if (n <= 32) {
                // lane a holds candidate a; rank = position in the reference's visiting order (context.cpp:214-242)
                const bool mine = lane < n;
                CInfo rec = {0, 0, 0, 0};
                uint32_t key = 0xffffffffu;
                if (mine) {
                    rec = s_u[lane];
                    key = rec.sortkey;
                }
                int rank = 0;
                const int n4 = (n + 3) >> 2;
                for (int u4 = 0; u4 < n4; u4++) {
                    const uint4 k4 = reinterpret_cast<const uint4*>(s_key)[u4];
                    rank += (k4.x < key) + (k4.y < key) + (k4.z < key) + (k4.w < key);
                }
                const int cy = (int16_t)(rec.cyx & 0xffff), cx = rec.cyx >> 16;
                const uint32_t off = (uint32_t)(2 * ((ap.OY - cy) * TS + (ap.OX - cx)));
                const uint32_t rbit = mine ? (1u << rank) : 0u;
                // tile t lists the candidates whose window can touch its 32 columns; bit r of the tile's mask <=>
                // the rank-r candidate is listed, so list position = number of listed candidates of smaller rank
#pragma unroll
                for (int t = 0; t < TPS; t++) {
                    const bool h = mine && ((unsigned)(cx - (sj0 + 32 * t - S)) <= (unsigned)(2 * S + 31));
                    const uint32_t mt = __reduce_or_sync(FSLIC_FULL, h ? rbit : 0u);
                    if (h) {
                        const int pos = __popc(mt & (rbit - 1u));
                        s_ent[t][pos] = make_uint2(rec.color, off);
                        s_tk[t][pos] = (uint16_t)(rec.sortkey & 0xffffu);
                    }
                    ncnt |= (uint32_t)__popc(mt) << (8 * t);
                }
            }
            __syncwarp();
        }
        const bool ovf = n > 32;  // clusters piled on one spot: brute force for this super tile

        // ---- the quad tile of this super tile has landed; the staging tile is free again ----
        a5_mbar_wait(bar, phase);
        phase ^= 1u;
        if (store_pending) {
            if (lane == 0) a5_store_wait_read();
            __syncwarp();
            store_pending = false;
        }

        bool slow = false;  // warp uniform: the super tile needs the per-pixel store path
#pragma unroll 1
        for (int tq = 0; tq < ntile; tq++) {
            const int n_t = (int)((ncnt >> (8 * tq)) & 0xffu);
            const int tj0 = sj0 + 32 * tq;
            const int j = tj0 + lane;
            const bool last = tq == ntile - 1;

            // ---- 1. pixels ----
            uint32_t q[R];
#pragma unroll
            for (int rr = 0; rr < R; rr++) q[rr] = s_quad[rr * BW + 32 * tq + lane];

            if (ovf) {
                // ---- overflow: brute force straight from the cell grid, direct stores and atomics ----
                const size_t img_off = (size_t)b * H * W;
#pragma unroll
                for (int rr = 0; rr < R; rr++) {
                    if (j < W && rr < nrow) {
                        const int i = wi0 + rr * stride;
                        const uint32_t label = assign_pixel_generic(ap, i, j, q[rr], ci, cs, labels + img_off);
                        if (UPDATE && label != 0xFFFF) acc_add_pixel(ac, label, i, j, q[rr]);
                    }
                }
                if (last) {
                    __syncwarp();
                    if (has_next && lane == 0) {
                        a5_mbar_expect_tx(bar, QBYTES);
                        a5_tma_load_3d(quad_dst, &tm_quad, nsx * BW, nty * R, nb, bar);
                    }
                }
                continue;
            }

            // ---- 2. distances ----
            // every (row, column) of the footprint is inside the patch for every listed candidate, valid or not.
            // patch entry of (row rr, candidate c) at shared byte address row0 + c.offset + rr * 2*stride*TS
            const unsigned char* rowp = smem_raw + 2 * (wi0 * TS + j);
            uint32_t best[R];
#pragma unroll
            for (int rr = 0; rr < R; rr++) best[rr] = 0xffffffffu;
            for (int c = 0; c < n_t; c++) {
                const uint2 e = s_ent[tq][c];
                const unsigned char* pc = rowp + (int)e.y;
#pragma unroll
                for (int rr = 0; rr < R; rr++) {
                    const uint32_t sp = *reinterpret_cast<const uint16_t*>(pc + rr * (2 * stride * TS));
                    const uint32_t d = sad4_acc(q[rr], e.x, sp);
                    best[rr] = min(best[rr], d * 65536u + (uint32_t)c);
                }
            }
            if (last) {
                // every lane has consumed its pixels of the last tile: the quad buffer may be overwritten
                __syncwarp();
                if (has_next && lane == 0) {
                    a5_mbar_expect_tx(bar, QBYTES);
                    a5_tma_load_3d(quad_dst, &tm_quad, nsx * BW, nty * R, nb, bar);
                }
            }

            // ---- 3. labels into the staging tile ----
            uint32_t rw = 0;  // local rank bytes, one per row (0x7f = contributes to no candidate)
            bool allcov = true;
            const uint32_t colbad = (edge && j >= W) ? 0xffffffffu : 0u;
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                uint32_t key = best[rr];
                if (edge) key |= colbad | ((rr < nrow) ? 0u : 0xffffffffu);
                const bool cov = key < A5_BIGKEY;
                allcov = allcov && cov;
                uint32_t rb = 0x7fu, lab = 0xFFFEu;  // 0xFFFE: "not covered" marker of the staging tile (K <= 65533)
                if (cov) {
                    rb = key & 0xffu;
                    lab = s_tk[tq][rb];
                }
                s_lab[rr * BW + 32 * tq + lane] = (uint16_t)lab;
                rw |= rb << (8 * rr);
            }
            if (!__all_sync(FSLIC_FULL, allcov)) slow = true;

            // ---- 4. update sums on the tensor cores (context.cpp:316-327) ----
            if (UPDATE) {
                // D[candidate][feature] += OneHot[candidate][pixel] * F[pixel][feature]   (m16n8k32, u8 x u8 -> s32)
                //   A = one-hot of the winning rank, built in registers (16 candidates per pass);
                //   B = [1, row, lane, L, a, b, 0, 0] per pixel, staged in shared memory.
                // Lane (g, tig) ends up with features (2 tig, 2 tig + 1) of candidates g and g + 8: exactly the two
                // halves of packed accumulator word tig.
                const uint32_t lo01 = __byte_perm(q[0], q[1], 0x5140), lo23 = __byte_perm(q[2], q[3], 0x5140);
                const uint32_t hi01 = __byte_perm(q[0], q[1], 0x0062), hi23 = __byte_perm(q[2], q[3], 0x0062);
                *reinterpret_cast<uint4*>(&s_feat[lane][0]) =
                    make_uint4(0x01010101u, 0x03020100u, (uint32_t)lane * 0x01010101u, __byte_perm(lo01, lo23, 0x5410));
                *reinterpret_cast<uint4*>(&s_feat[lane][4]) =
                    make_uint4(__byte_perm(lo01, lo23, 0x7632), __byte_perm(hi01, hi23, 0x5410), 0u, 0u);
                __syncwarp();
                const int n16 = (n_t + 15) >> 4;
                for (int nt = 0; nt < n16; nt++) {
                    int d[4] = {0, 0, 0, 0};
                    const uint32_t mg0 = (uint32_t)(nt * 16 + g) * 0x01010101u, mg1 = mg0 + 0x08080808u;
                    const bool two = n_t > nt * 16 + 8;  // candidates g + 8 exist in this pass
                    if (two) {
#pragma unroll
                        for (int s4 = 0; s4 < 4; s4++) {
                            const uint32_t w0 = __shfl_sync(FSLIC_FULL, rw, 8 * s4 + tig);
                            const uint32_t w1 = __shfl_sync(FSLIC_FULL, rw, 8 * s4 + 4 + tig);
                            mma_u8_16x8x32(d, eq7(w0, mg0), eq7(w0, mg1), eq7(w1, mg0), eq7(w1, mg1),
                                           s_feat[8 * s4 + tig][g], s_feat[8 * s4 + 4 + tig][g]);
                        }
                    } else {
#pragma unroll
                        for (int s4 = 0; s4 < 4; s4++) {
                            const uint32_t w0 = __shfl_sync(FSLIC_FULL, rw, 8 * s4 + tig);
                            const uint32_t w1 = __shfl_sync(FSLIC_FULL, rw, 8 * s4 + 4 + tig);
                            mma_u8_16x8x32(d, eq7(w0, mg0), 0u, eq7(w1, mg0), 0u, s_feat[8 * s4 + tig][g],
                                           s_feat[8 * s4 + 4 + tig][g]);
                        }
                    }
                    // sums are scaled by 128 (the one-hot byte is 0x80)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        if (hh == 1 && !two) break;
                        const int c = nt * 16 + g + 8 * hh;
                        const uint32_t v0 = (uint32_t)d[2 * hh] >> 7, v1 = (uint32_t)d[2 * hh + 1] >> 7;
                        const uint32_t cnt = __shfl_sync(FSLIC_FULL, v0, lane & ~3);  // feature 0 lives in the tig = 0 lane
                        if (c < n_t && tig < 3 && cnt != 0) {
                            unsigned long long word;
                            if (tig == 0)
                                word = (unsigned long long)cnt |
                                       ((unsigned long long)(cnt * (uint32_t)wi0 + (uint32_t)stride * v1) << 32);
                            else if (tig == 1)
                                word = (unsigned long long)(cnt * (uint32_t)tj0 + v0) | ((unsigned long long)v1 << 32);
                            else
                                word = (unsigned long long)v0 | ((unsigned long long)v1 << 32);
                            atomicAdd(&ac[(uint32_t)s_tk[tq][c] * 4 + tig], word);
                        }
                    }
                }
                __syncwarp();  // s_feat is rewritten by the next tile / the next list
            }
        }

        // ---- 5. labels out ----
        if (!ovf) {
            if (!slow) {
                a5_fence_async();  // the staging tile was written through the generic proxy
                __syncwarp();
                if (lane == 0) a5_tma_store_3d(&tm_lab, lab_src, sj0, wsr0, b);
                store_pending = true;
            } else {
                // per-pixel path: the image edge clips this super tile, or some pixel is covered by no window
                __syncwarp();
                const size_t img_off = (size_t)b * H * W;
                for (int tq = 0; tq < ntile; tq++) {
                    const int j = sj0 + 32 * tq + lane;
                    if (j >= W) continue;
                    for (int rr = 0; rr < nrow; rr++) {
                        const int i = wi0 + rr * stride;
                        const uint32_t v = s_lab[rr * BW + 32 * tq + lane];
                        uint16_t* lp = labels + img_off + (size_t)i * W + j;
                        if (v != 0xFFFEu) {
                            *lp = (uint16_t)v;
                        } else if ((i % ap.cfg_stride) >= ap.fresh_from) {
                            *lp = 0xFFFF;  // never assigned before: the reference's map still holds 0xFFFF here
                        } else if (UPDATE) {  // a stale label from an earlier pass still counts (context.cpp:318-319)
                            const uint16_t old = *lp;
                            if (old != 0xFFFF) acc_add_pixel(ac, old, i, j, quad[img_off + (size_t)i * W + j]);
                        }
                    }
                }
            }
        }
        __syncwarp();  // list staging, staging tile and MMA staging are rewritten by the next super tile
        b = nb; ty = nty; sx = nsx;
    }
    if (store_pending && lane == 0) a5_store_wait_read();  // the staging tile must outlive the last bulk store's read
}
