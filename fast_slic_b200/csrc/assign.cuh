// fast_slic_b200/csrc/assign.cuh -- the assign + update hot loop.
//
// Replaces BaseContext::assign / assign_clusters / update of the reference
// (/root/reference/src/context.cpp:200-243, 259-298, 302-387; AVX2 form arch/x64/avx2.h:11-185).
//
// The reference is cluster-centric: every cluster scatters into its (2S+1)^2 window with a strict
// `<` against a per-pixel running minimum; ties are therefore won by the cluster visited first,
// i.e. by the smaller (phase, k) where phase = 2*((cy/T)&1) + ((cx/T)&1), T = 2S+32
// (context.cpp:214-242).  On the GPU the loop is turned inside out: every pixel gathers over the
// clusters whose window covers it and minimises the packed key  d << 16 | rank,  rank being the
// candidate's position in the (phase, k)-sorted candidate list of its CTA tile.  Pixels no window
// covers keep their previous label (context.cpp:289-294 never fires for them).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// k_prepare: per-image bookkeeping between two assign passes (one CTA per image).
//   1. finalise the previous update: integer round-divide of the accumulated sums
//      (context.cpp:356-374, round_int fast-slic-common.h:63-65) and clear the accumulators;
//   2. first pass only: re-seed the cluster colour from the quad image (context.cpp:128-135);
//   3. clamp centres into the image (context.cpp:209-212), truncate to int16 (context.cpp:266-267),
//      derive the visiting-order key, write the 16-byte CInfo record;
//   4. counting-sort the clusters into a uniform cell grid of pitch G >= S, so a tile can collect
//      the clusters whose window may touch it from a few contiguous ranges.
// ---------------------------------------------------------------------------------------------
struct PrepParams {
    int H, W, K, S, T;
    int G, cellW, cellH, ncell;
    int first;     // 1: re-seed colours, no update to finalise
    int finalize;  // 1: fold `acc` into the clusters
    int last;      // 1: after the final update: set is_active / is_updatable like the reference leaves them
};

__global__ void __launch_bounds__(1024) k_prepare(PrepParams pp, fslic_cluster* __restrict__ clusters,
                                                   uint32_t* __restrict__ acc, const uint32_t* __restrict__ quad,
                                                   CInfo* __restrict__ cinfo, int* __restrict__ cell_start,
                                                   int* __restrict__ cell_items) {
    extern __shared__ int s_cnt[];  // ncell + 1 counters, then 1024/32 warp sums
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    fslic_cluster* cl = clusters + (size_t)b * pp.K;
    uint32_t* ac = acc + (size_t)b * pp.K * 6;
    const uint32_t* qd = quad + (size_t)b * pp.H * pp.W;
    CInfo* ci = cinfo + (size_t)b * pp.K;
    int* cs = cell_start + (size_t)b * (pp.ncell + 1);
    int* items = cell_items + (size_t)b * pp.K;

    for (int c = tid; c <= pp.ncell; c += nt) s_cnt[c] = 0;
    __syncthreads();

    for (int k = tid; k < pp.K; k += nt) {
        fslic_cluster c = cl[k];
        if (pp.finalize) {
            const uint32_t n = ac[k * 6 + 0];
            c.num_members = n;  // written even when n == 0 (context.cpp:360-362)
            if (n > 0) {
                const int32_t in = (int32_t)n, half = in / 2;
                c.y = (float)(((int32_t)ac[k * 6 + 1] + half) / in);
                c.x = (float)(((int32_t)ac[k * 6 + 2] + half) / in);
                c.r = (float)(((int32_t)ac[k * 6 + 3] + half) / in);
                c.g = (float)(((int32_t)ac[k * 6 + 4] + half) / in);
                c.b = (float)(((int32_t)ac[k * 6 + 5] + half) / in);
            }
#pragma unroll
            for (int f = 0; f < 6; f++) ac[k * 6 + f] = 0;
        }
        if (pp.first) {
            int y = min(max((int)c.y, 0), pp.H - 1), x = min(max((int)c.x, 0), pp.W - 1);
            const uint32_t q = qd[(size_t)y * pp.W + x];
            c.r = (float)(q & 0xff);
            c.g = (float)((q >> 8) & 0xff);
            c.b = (float)((q >> 16) & 0xff);
        }
        // safeguard clamp, stored back like the reference does
        c.x = fminf(fmaxf(c.x, 0.f), (float)(pp.W - 1));
        c.y = fminf(fmaxf(c.y, 0.f), (float)(pp.H - 1));
        c.number = (uint16_t)k;
        c.is_active = 1;
        c.is_updatable = 2;
        cl[k] = c;

        const int cy = (int16_t)c.y, cx = (int16_t)c.x;
        const int cr = (int16_t)c.r, cg = (int16_t)c.g, cb = (int16_t)c.b;
        const int phase = 2 * ((cy / pp.T) & 1) + ((cx / pp.T) & 1);
        CInfo r;
        r.cyx = (cy & 0xffff) | (cx << 16);
        r.color = (uint32_t)(cr & 0xff) | ((uint32_t)(cg & 0xff) << 8) | ((uint32_t)(cb & 0xff) << 16);
        r.sortkey = ((uint32_t)phase << 16) | (uint32_t)k;
        r.pad = 0;
        ci[k] = r;
        const int cell = (cy / pp.G) * pp.cellW + (cx / pp.G);
        atomicAdd(&s_cnt[cell], 1);
    }
    __syncthreads();
    // exclusive scan of the cell histogram (ncell <= 16384): serial over chunks of blockDim
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base <= pp.ncell; base += nt) {
        const int c = base + tid;
        const int v = (c <= pp.ncell) ? s_cnt[c] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FSLIC_FULL, x, o);
            if ((tid & 31) >= o) x += y;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            int w = (tid < (nt >> 5)) ? s_warp[tid] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(FSLIC_FULL, w, o);
                if (tid >= o) w += y;
            }
            s_warp[tid] = w;
        }
        __syncthreads();
        const int woff = (tid >> 5) ? s_warp[(tid >> 5) - 1] : 0;
        const int excl = s_carry + woff + x - v;
        __syncthreads();
        if (c <= pp.ncell) {
            s_cnt[c] = excl;  // becomes the running fill pointer
            cs[c] = excl;
        }
        if (tid == nt - 1) s_carry = excl + v;
        __syncthreads();
    }
    for (int k = tid; k < pp.K; k += nt) {
        const int32_t cyx = ci[k].cyx;
        const int cy = (int16_t)(cyx & 0xffff), cx = cyx >> 16;
        const int cell = (cy / pp.G) * pp.cellW + (cx / pp.G);
        const int slot = atomicAdd(&s_cnt[cell], 1);
        items[slot] = k;
    }
}

// ---------------------------------------------------------------------------------------------
// Spatial patch (BaseContext::set_spatial_patch, context.cpp:23-40), laid out for LINEAR addressing:
//   tbl[(di + OY) * TS + (dj + OX)] = (u16)(coef * (float)(|di| + |dj|))  inside the (2S+1)^2 window,
//                                   = FSLIC_BIGSP                         outside it,
// for di in [-OY, OY], dj in [-OX, OX].  A pixel's entry for candidate c is then at
//   (i*TS + j) + ((OY - cy)*TS + (OX - cx)):  a per-thread constant plus a per-candidate constant,
// so the window predicate and both abs() disappear from the inner loop.
// ---------------------------------------------------------------------------------------------
__global__ void k_build_sptable(uint16_t* __restrict__ tbl, int S, int OY, int OX, int TS, float coef) {
    const int n = (2 * OY + 1) * TS;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const int r = t / TS, c = t - r * TS;
        const int di = abs(r - OY), dj = abs(c - OX);
        uint16_t v = (uint16_t)FSLIC_BIGSP;
        if (di <= S && dj <= S && c <= 2 * OX) v = (uint16_t)__float2uint_rz(__fmul_rn(coef, (float)(di + dj)));
        tbl[t] = v;
    }
}

struct AssignParams {
    int H, W, K, S, B;
    int stride, rem;   // rows i with i % stride == rem are processed; sub-row sr <-> i = rem + sr*stride
    int nsub;          // number of such rows
    int cfg_stride;    // the configured subsample stride (freshness test)
    int fresh_from;    // rows with (i % cfg_stride) >= fresh_from were never assigned before
    int G, cellW, cellH, ncell;
    int OY, OX, TS, tbl_elems;
    int tiles_x, tiles_y, ntiles;  // per image
    float coef;        // generic path only
};

#define AS_WARPS 16
#define AS_THREADS (AS_WARPS * 32)
#define AS_WX 4            // warps across  -> tile is 128 columns
#define AS_WY 4            // warps down
#define AS_NC 256          // candidate capacity of a tile; beyond it the tile goes to the generic kernel

// ---------------------------------------------------------------------------------------------
// k_assign_tiles<R, UPDATE>: persistent CTAs; each loops over tiles of 128 columns x (4*R) sub-rows.
//   per tile : collect the clusters whose window can touch the tile from the cell grid, sort them
//              by (phase, k) (rank-by-counting), keep them in shared memory;
//   per warp : 32 columns x R sub-rows; ballot-filter the tile list down to the clusters that can
//              touch the warp's footprint (order preserved);
//   per pixel and candidate:  LDS.U16 patch entry -> VABSDIFF4.U8.ACC (colour SAD + spatial) ->
//              IMAD (d << 16 | rank) -> VIMNMX;
//   update   : warp-aggregated (ballot + REDUX) sums into per-tile shared accumulators, flushed with
//              one RED per touched (cluster, field) -- exact, integer sums are order independent.
// HBM per processed pixel: 4 B quad read + 2 B label written.
// ---------------------------------------------------------------------------------------------
template <int R, bool UPDATE>
__global__ void __launch_bounds__(AS_THREADS) k_assign_tiles(AssignParams ap, const uint32_t* __restrict__ quad,
                                                              uint16_t* __restrict__ labels,
                                                              const CInfo* __restrict__ cinfo,
                                                              const int* __restrict__ cell_start,
                                                              const int* __restrict__ cell_items,
                                                              uint32_t* __restrict__ acc,
                                                              const uint16_t* __restrict__ g_tbl,
                                                              int* __restrict__ overflow_list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t* s_tbl = reinterpret_cast<uint16_t*>(smem_raw);
    __shared__ uint2 s_ent[AS_NC];      // sorted: {colour, patch base offset (u16 elements)}
    __shared__ int32_t s_cyx[AS_NC];    // sorted: packed centre, for the warp filter
    __shared__ uint16_t s_k[AS_NC];     // sorted: cluster index
    __shared__ uint32_t s_ukey[AS_NC];  // unsorted gather
    __shared__ uint32_t s_ucol[AS_NC];
    __shared__ int32_t s_ucyx[AS_NC];
    __shared__ uint32_t s_acc[UPDATE ? AS_NC * 6 : 1];
    __shared__ int s_n;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wx = warp % AS_WX, wy = warp / AS_WX;

    // the patch is loaded once per CTA and reused for every tile this CTA processes
    for (int t = tid; t < (ap.tbl_elems + 1) / 2; t += AS_THREADS)
        reinterpret_cast<uint32_t*>(s_tbl)[t] = reinterpret_cast<const uint32_t*>(g_tbl)[t];

    const int S = ap.S, W = ap.W, H = ap.H;
    const long total_tiles = (long)ap.ntiles * ap.B;
    for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = (int)(tile / ap.ntiles);
        const int tl = (int)(tile - (long)b * ap.ntiles);
        const int ty = tl / ap.tiles_x, tx = tl - ty * ap.tiles_x;
        const int tj0 = tx * (AS_WX * 32);
        const int tsr0 = ty * (AS_WY * R);
        const int tsr1 = min(tsr0 + AS_WY * R, ap.nsub) - 1;  // last valid sub-row of the tile
        const int ti0 = ap.rem + tsr0 * ap.stride, ti1 = ap.rem + tsr1 * ap.stride;
        const int tj1 = min(tj0 + AS_WX * 32, W) - 1;

        const CInfo* ci = cinfo + (size_t)b * ap.K;
        const int* cs = cell_start + (size_t)b * (ap.ncell + 1);
        const int* items = cell_items + (size_t)b * ap.K;

        __syncthreads();  // previous tile fully consumed (also orders the table load on the first pass)
        if (tid == 0) s_n = 0;
        if (UPDATE)
            for (int t = tid; t < AS_NC * 6; t += AS_THREADS) s_acc[t] = 0;
        __syncthreads();

        // ---- gather: clusters with cy in [ti0-S, ti1+S], cx in [tj0-S, tj1+S] ----
        {
            const int cr0 = max(ti0 - S, 0) / ap.G, cr1 = min(ti1 + S, H - 1) / ap.G;
            const int cc0 = max(tj0 - S, 0) / ap.G, cc1 = min(tj1 + S, W - 1) / ap.G;
            for (int cr = cr0 + warp; cr <= cr1; cr += AS_WARPS) {
                const int s = cs[cr * ap.cellW + cc0], e = cs[cr * ap.cellW + cc1 + 1];
                for (int t = s + lane; t < e; t += 32) {
                    const int k = items[t];
                    const CInfo r = ci[k];
                    const int cy = (int16_t)(r.cyx & 0xffff), cx = r.cyx >> 16;
                    if (cy >= ti0 - S && cy <= ti1 + S && cx >= tj0 - S && cx <= tj1 + S) {
                        const int slot = atomicAdd(&s_n, 1);
                        if (slot < AS_NC) {
                            s_ukey[slot] = r.sortkey;
                            s_ucol[slot] = r.color;
                            s_ucyx[slot] = r.cyx;
                        }
                    }
                }
            }
        }
        __syncthreads();
        const int n = s_n;
        if (n > AS_NC) {  // too many clusters around this tile: defer to the generic kernel
            if (tid == 0) {
                const int slot = atomicAdd(&overflow_list[0], 1);
                overflow_list[1 + slot] = (int)tile;
            }
            continue;
        }
        // ---- sort by (phase, k): rank by counting (keys are unique) ----
        for (int t = tid; t < n; t += AS_THREADS) {
            const uint32_t key = s_ukey[t];
            int rank = 0;
            for (int u = 0; u < n; u++) rank += (s_ukey[u] < key);
            const int32_t cyx = s_ucyx[t];
            const int cy = (int16_t)(cyx & 0xffff), cx = cyx >> 16;
            s_ent[rank] = make_uint2(s_ucol[t], (uint32_t)((ap.OY - cy) * ap.TS + (ap.OX - cx)));
            s_cyx[rank] = cyx;
            s_k[rank] = (uint16_t)(key & 0xffff);
        }
        __syncthreads();

        // ---- per-warp footprint ----
        const int wj0 = tj0 + wx * 32;
        const int wsr0 = tsr0 + wy * R;
        if (wj0 < W && wsr0 < ap.nsub) {  // warp-uniform
            const int wsr1 = min(wsr0 + R, ap.nsub) - 1;
            const int wi0 = ap.rem + wsr0 * ap.stride, wi1 = ap.rem + wsr1 * ap.stride;
            const int j = wj0 + lane;
            const bool colok = j < W;
            const uint32_t* qd = quad + (size_t)b * H * W;
            uint16_t* lb = labels + (size_t)b * H * W;

            uint32_t q[R], best[R];
            int pixoff[R];
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                const int i = wi0 + rr * ap.stride;
                const bool ok = colok && (wsr0 + rr < ap.nsub);
                q[rr] = ok ? ld_nc_u32(qd + (size_t)i * W + j) : 0u;
                // invalid lanes alias the warp origin so every patch address stays in range
                pixoff[rr] = ok ? (i * ap.TS + j) : (wi0 * ap.TS + wj0);
                best[rr] = 0xffffffffu;
            }

            for (int base = 0; base < n; base += 32) {
                const int c = base + lane;
                bool hit = false;
                if (c < n) {
                    const int32_t cyx = s_cyx[c];
                    const int cy = (int16_t)(cyx & 0xffff), cx = cyx >> 16;
                    hit = (cy >= wi0 - S) && (cy <= wi1 + S) && (cx >= wj0 - S) && (cx <= wj0 + 31 + S);
                }
                unsigned m = __ballot_sync(FSLIC_FULL, hit);
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    const int cidx = base + bit;
                    const uint2 e = s_ent[cidx];
#pragma unroll
                    for (int rr = 0; rr < R; rr++) {
                        const uint32_t sp = s_tbl[(int)e.y + pixoff[rr]];
                        const uint32_t d = sad4_acc(q[rr], e.x, sp);
                        best[rr] = min(best[rr], d * 65536u + (uint32_t)cidx);
                    }
                }
            }

            // ---- labels + update ----
            uint32_t* ac = acc + (size_t)b * ap.K * 6;
#pragma unroll
            for (int rr = 0; rr < R; rr++) {
                const int i = wi0 + rr * ap.stride;
                const bool ok = colok && (wsr0 + rr < ap.nsub);
                const bool covered = ok && ((best[rr] >> 16) < FSLIC_BIGSP);
                const int rank = (int)(best[rr] & 0xffff);
                if (covered) {
                    lb[(size_t)i * W + j] = s_k[rank];
                } else if (ok) {
                    const bool fresh = (i % ap.cfg_stride) >= ap.fresh_from;
                    if (fresh) {
                        lb[(size_t)i * W + j] = 0xFFFF;
                    } else if (UPDATE) {  // a stale label from an earlier pass still counts (context.cpp:318-319)
                        const uint16_t old = lb[(size_t)i * W + j];
                        if (old != 0xFFFF) {
                            atomicAdd(&ac[old * 6 + 0], 1u);
                            atomicAdd(&ac[old * 6 + 1], (uint32_t)i);
                            atomicAdd(&ac[old * 6 + 2], (uint32_t)j);
                            atomicAdd(&ac[old * 6 + 3], q[rr] & 0xff);
                            atomicAdd(&ac[old * 6 + 4], (q[rr] >> 8) & 0xff);
                            atomicAdd(&ac[old * 6 + 5], (q[rr] >> 16) & 0xff);
                        }
                    }
                }
                if (UPDATE) {
                    unsigned act = __ballot_sync(FSLIC_FULL, covered);
                    const uint32_t w0 = (q[rr] & 0xff) | ((q[rr] & 0xff00) << 8);        // L | a << 16
                    const uint32_t w1 = ((q[rr] >> 16) & 0xff) | ((uint32_t)lane << 16);  // b | lane << 16
                    while (act) {
                        const int src = __ffs(act) - 1;
                        const int r0 = __shfl_sync(FSLIC_FULL, rank, src);
                        const bool mine = covered && rank == r0;
                        const unsigned mm = __ballot_sync(FSLIC_FULL, mine);
                        if (mine) {
                            const uint32_t s0 = __reduce_add_sync(mm, w0);
                            const uint32_t s1 = __reduce_add_sync(mm, w1);
                            if (lane == src) {
                                const uint32_t cnt = __popc(mm);
                                uint32_t* sa = &s_acc[r0 * 6];
                                atomicAdd(&sa[0], cnt);
                                atomicAdd(&sa[1], cnt * (uint32_t)i);
                                atomicAdd(&sa[2], cnt * (uint32_t)wj0 + (s1 >> 16));
                                atomicAdd(&sa[3], s0 & 0xffff);
                                atomicAdd(&sa[4], s0 >> 16);
                                atomicAdd(&sa[5], s1 & 0xffff);
                            }
                        }
                        act &= ~mm;
                    }
                }
            }
        }
        if (UPDATE) {
            __syncthreads();
            uint32_t* ac = acc + (size_t)b * ap.K * 6;
            for (int t = tid; t < n * 6; t += AS_THREADS) {
                const uint32_t v = s_acc[t];
                if (v) atomicAdd(&ac[(int)s_k[t / 6] * 6 + (t % 6)], v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_assign_generic<UPDATE>: correctness-first path without the shared-memory patch or tile lists.
// Used (a) for whole passes when S is so large that the linear patch does not fit in shared memory,
// (b) for the few tiles whose candidate list overflowed AS_NC (clusters piled on one spot).
// One thread per pixel; walks the cell grid; lexicographic minimum of (d, phase, k) in a u64 key.
// ---------------------------------------------------------------------------------------------
template <bool UPDATE>
__global__ void __launch_bounds__(256) k_assign_generic(AssignParams ap, const uint32_t* __restrict__ quad,
                                                         uint16_t* __restrict__ labels,
                                                         const CInfo* __restrict__ cinfo,
                                                         const int* __restrict__ cell_start,
                                                         const int* __restrict__ cell_items,
                                                         uint32_t* __restrict__ acc,
                                                         const int* __restrict__ tile_list, int R) {
    // tile_list == nullptr: blockIdx.x enumerates all tiles of all images; else tile_list[0] = count
    const long count = tile_list ? (long)tile_list[0] : (long)ap.ntiles * ap.B;
    for (long idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const long tile = tile_list ? (long)tile_list[1 + idx] : idx;
    const int b = (int)(tile / ap.ntiles);
    const int tl = (int)(tile - (long)b * ap.ntiles);
    const int ty = tl / ap.tiles_x, tx = tl - ty * ap.tiles_x;
    const int tj0 = tx * (AS_WX * 32), tsr0 = ty * (AS_WY * R);
    const int nrows = AS_WY * R, ncols = AS_WX * 32;
    const int S = ap.S, W = ap.W, H = ap.H;
    const CInfo* ci = cinfo + (size_t)b * ap.K;
    const int* cs = cell_start + (size_t)b * (ap.ncell + 1);
    const int* items = cell_items + (size_t)b * ap.K;
    const uint32_t* qd = quad + (size_t)b * H * W;
    uint16_t* lb = labels + (size_t)b * H * W;
    uint32_t* ac = acc + (size_t)b * ap.K * 6;
    for (int t = threadIdx.x; t < nrows * ncols; t += blockDim.x) {
        const int sr = tsr0 + t / ncols, j = tj0 + t % ncols;
        if (sr >= ap.nsub || j >= W) continue;
        const int i = ap.rem + sr * ap.stride;
        const uint32_t q = qd[(size_t)i * W + j];
        unsigned long long best = ~0ull;
        const int cr0 = max(i - S, 0) / ap.G, cr1 = min(i + S, H - 1) / ap.G;
        const int cc0 = max(j - S, 0) / ap.G, cc1 = min(j + S, W - 1) / ap.G;
        for (int cr = cr0; cr <= cr1; cr++) {
            const int s = cs[cr * ap.cellW + cc0], e = cs[cr * ap.cellW + cc1 + 1];
            for (int u = s; u < e; u++) {
                const CInfo r = ci[items[u]];
                const int cy = (int16_t)(r.cyx & 0xffff), cx = r.cyx >> 16;
                const int di = abs(i - cy), dj = abs(j - cx);
                if (di > S || dj > S) continue;
                const uint32_t sp = (uint16_t)__float2uint_rz(__fmul_rn(ap.coef, (float)(di + dj)));
                const uint32_t d = sad4_acc(q, r.color, sp) & 0xffffu;  // u16 arithmetic like the scalar reference
                const unsigned long long key = ((unsigned long long)d << 32) | r.sortkey;
                best = key < best ? key : best;
            }
        }
        uint32_t label;
        if (best != ~0ull && (uint32_t)(best >> 32) < 0xFFFFu) {
            label = (uint32_t)(best & 0xffff);
            lb[(size_t)i * W + j] = (uint16_t)label;
        } else {
            const bool fresh = (i % ap.cfg_stride) >= ap.fresh_from;
            if (fresh) {
                lb[(size_t)i * W + j] = 0xFFFF;
                label = 0xFFFF;
            } else {
                label = lb[(size_t)i * W + j];
            }
        }
        if (UPDATE && label != 0xFFFF) {
            atomicAdd(&ac[label * 6 + 0], 1u);
            atomicAdd(&ac[label * 6 + 1], (uint32_t)i);
            atomicAdd(&ac[label * 6 + 2], (uint32_t)j);
            atomicAdd(&ac[label * 6 + 3], q & 0xff);
            atomicAdd(&ac[label * 6 + 4], (q >> 8) & 0xff);
            atomicAdd(&ac[label * 6 + 5], (q >> 16) & 0xff);
        }
    }
    }
}
