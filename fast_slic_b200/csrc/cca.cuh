// fast_slic_b200/csrc/cca.cuh -- connectivity enforcement on the GPU.
//
// Replaces cca::ConnectivityEnforcer::execute with assign_disjoint_set / DisjointSet::merge /
// DisjointSet::flatten (/root/reference/src/cca.cpp:33-101, 103-173, 178-265; cca.h:36-57).
//
// Pipeline (all kernels batched over images with blockIdx.y):
//   k_ccl_init     parent[p] = first pixel of p's run inside its 32-pixel chunk
//   k_ccl_merge    lock-free union (atomicMin on roots) across chunk seams and between rows; the
//                  representative of a set is always its MINIMUM raster index == the reference's
//                  "leader" (cca.h:38-55 merges towards the smaller index)
//   k_ccl_flatten  parent[p] = root; per-root area (run-aggregated atomics); roots per block
//   k_scan_blocks  exclusive scan of the per-block counts (one CTA per image)
//   k_ccl_number   component number = rank of the root in raster order (cca.cpp:118-134);
//                  scatters leader / area by component number; counts candidates area >= thres
//   k_cca_select   only when candidates > K: libstdc++ std::partial_sort set semantics
//                  (cca.cpp:225-228), emulated step for step (make_heap / adjust_heap / push_heap)
//   k_kept_count / k_scan_blocks / k_kept_label   new label = rank among kept components, which
//                  are already in leader order (cca.cpp:229-237)
//   k_cca_absorb   unkept components take the label of the component left of (or above) their
//                  leader, transitively (cca.cpp:238-255)
//   k_cca_output   out[p] = final label of root(p)  (cca.cpp:260-263)
#pragma once
#include "common.cuh"

#define CCA_BLOCK 1024

struct CcaParams {
    int H, W, N;       // N = H*W
    int K;             // max_label_size (cca.cpp:176)
    int thres;         // min_threshold
    int nblk;          // ceil(N / CCA_BLOCK)
    int heap_in_smem;  // k_cca_select keeps its heap in shared memory
};

// Per-image scalar scratch
struct CcaCounters {
    int ncomp;
    int ncand;
    int nkept;
    int sel_mode;  // 1: selection ran, kept flag = top bit of carea
};

__device__ __forceinline__ int ccl_find(const int* par, int x) {
    int p = par[x];
    while (p != x) {
        x = p;
        p = par[x];
    }
    return x;
}

// lock-free union keeping the smaller index as root (Komura-style atomicMin reduction)
__device__ __forceinline__ void ccl_union(int* par, int a, int b) {
    a = ccl_find(par, a);
    b = ccl_find(par, b);
    while (a != b) {
        if (a < b) {
            int t = a;
            a = b;
            b = t;
        }  // a > b: hang a under b
        const int old = atomicMin(&par[a], b);
        if (old == a) break;  // a was still a root: done
        a = ccl_find(par, old);  // someone re-parented a meanwhile: continue from there
        b = ccl_find(par, b);
    }
}

__global__ void __launch_bounds__(CCA_BLOCK) k_ccl_init(CcaParams cp, const uint16_t* __restrict__ labels,
                                                        int* __restrict__ par, uint32_t* __restrict__ area_at) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * CCA_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint16_t* lab = labels + (size_t)b * cp.N;
    const bool ok = p < cp.N;
    const uint32_t v = ok ? lab[p] : 0x10000u;
    const uint32_t left = __shfl_up_sync(FSLIC_FULL, v, 1);
    const int j = ok ? (p % cp.W) : 0;
    const bool start = (lane == 0) || (j == 0) || (v != left);
    const unsigned m = __ballot_sync(FSLIC_FULL, start);
    const int s = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
    if (ok) {
        par[(size_t)b * cp.N + p] = p - (lane - s);
        area_at[(size_t)b * cp.N + p] = 0;
    }
}

__global__ void __launch_bounds__(CCA_BLOCK) k_ccl_merge(CcaParams cp, const uint16_t* __restrict__ labels,
                                                         int* __restrict__ par_all) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * CCA_BLOCK + threadIdx.x;
    if (p >= cp.N) return;
    const int lane = threadIdx.x & 31;
    const uint16_t* lab = labels + (size_t)b * cp.N;
    int* par = par_all + (size_t)b * cp.N;
    const int W = cp.W;
    const int j = p % W;
    const uint16_t v = lab[p];
    const bool has_left = j > 0, has_up = p >= W;
    const uint16_t left = has_left ? lab[p - 1] : 0;
    // seam between two 32-pixel chunks of the same row
    if (lane == 0 && has_left && left == v) ccl_union(par, p - 1, p);
    if (has_up) {
        const uint16_t up = lab[p - W];
        if (up == v) {
            // the pair (p, up) is implied by (p-1, up-1) when both runs extend to the left
            bool need = !has_left || left != v;
            if (!need) need = lab[p - W - 1] != up;
            if (need) ccl_union(par, p - W, p);
        }
    }
}

__global__ void __launch_bounds__(CCA_BLOCK) k_ccl_flatten(CcaParams cp, const uint16_t* __restrict__ labels,
                                                           int* __restrict__ par_all,
                                                           uint32_t* __restrict__ area_all,
                                                           int* __restrict__ blkcnt) {
    __shared__ int s_cnt;
    const int b = blockIdx.y;
    const int p = blockIdx.x * CCA_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint16_t* lab = labels + (size_t)b * cp.N;
    int* par = par_all + (size_t)b * cp.N;
    const bool ok = p < cp.N;
    const uint32_t v = ok ? lab[p] : 0x10000u;
    const uint32_t left = __shfl_up_sync(FSLIC_FULL, v, 1);
    const int j = ok ? (p % cp.W) : 0;
    const bool start = (lane == 0) || (j == 0) || (v != left);
    const unsigned m = __ballot_sync(FSLIC_FULL, start);
    const int s = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
    int root = 0;
    if (ok && start) root = ccl_find(par, p);
    root = __shfl_sync(FSLIC_FULL, root, s);
    bool isroot = false;
    if (ok) {
        par[p] = root;
        if (start) {
            // run length: distance to the next run start (or the end of the chunk / image)
            const unsigned above = (lane == 31) ? 0u : (m >> (lane + 1));
            int len = above ? (__ffs(above)) : (32 - lane);
            if (p + len > cp.N) len = cp.N - p;
            atomicAdd(&area_all[(size_t)b * cp.N + root], (uint32_t)len);
            isroot = (root == p);
        }
    }
    const unsigned rm = __ballot_sync(FSLIC_FULL, isroot);
    if (lane == 0 && rm) atomicAdd(&s_cnt, __popc(rm));
    __syncthreads();
    if (threadIdx.x == 0) blkcnt[(size_t)b * cp.nblk + blockIdx.x] = s_cnt;
}

// exclusive scan of cnt[0..n) (n = *n_dev if n_dev else n_static), total -> *total_out
__global__ void __launch_bounds__(1024) k_scan_blocks(const int* __restrict__ cnt_all, int* __restrict__ off_all,
                                                      int stride_per_image, int n_static,
                                                      const int* __restrict__ n_dev_base, int n_dev_stride_ints,
                                                      int n_div, int* __restrict__ total_base,
                                                      int total_stride_ints) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int* cnt = cnt_all + (size_t)b * stride_per_image;
    int* off = off_all + (size_t)b * stride_per_image;
    int n = n_static;
    if (n_dev_base) n = (n_dev_base[(size_t)b * n_dev_stride_ints] + n_div - 1) / n_div;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int c = base + tid;
        const int v = (c < n) ? cnt[c] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FSLIC_FULL, x, o);
            if ((tid & 31) >= o) x += y;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            int w = s_warp[tid];
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
        if (c < n) off[c] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) total_base[(size_t)b * total_stride_ints] = s_carry;
}

__global__ void __launch_bounds__(CCA_BLOCK) k_ccl_number(CcaParams cp, const int* __restrict__ par_all,
                                                          uint32_t* __restrict__ aux_all,
                                                          const int* __restrict__ blkoff,
                                                          int* __restrict__ cleader_all,
                                                          uint32_t* __restrict__ carea_all,
                                                          CcaCounters* __restrict__ counters) {
    __shared__ int s_warp[32];
    __shared__ int s_cand;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x * CCA_BLOCK + tid;
    const int* par = par_all + (size_t)b * cp.N;
    uint32_t* aux = aux_all + (size_t)b * cp.N;
    if (tid == 0) s_cand = 0;
    const bool isroot = (p < cp.N) && (par[p] == p);
    const unsigned rm = __ballot_sync(FSLIC_FULL, isroot);
    if (lane == 0) s_warp[warp] = __popc(rm);
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FSLIC_FULL, x, o);
            if (lane >= o) x += y;
        }
        s_warp[lane] = x - w;
    }
    __syncthreads();
    bool cand = false;
    if (isroot) {
        const int c = blkoff[(size_t)b * cp.nblk + blockIdx.x] + s_warp[warp] + __popc(rm & ((1u << lane) - 1));
        const uint32_t a = aux[p];
        aux[p] = (uint32_t)c;  // the area slot of a root now holds its component number
        cleader_all[(size_t)b * cp.N + c] = p;
        carea_all[(size_t)b * cp.N + c] = a;
        cand = (int)a >= cp.thres;
    }
    const unsigned cm = __ballot_sync(FSLIC_FULL, cand);
    if (lane == 0 && cm) atomicAdd(&s_cand, __popc(cm));
    __syncthreads();
    if (tid == 0 && s_cand) atomicAdd(&counters[b].ncand, s_cand);
}

// ---------------------------------------------------------------------------------------------
// std::partial_sort(comps.begin(), comps.begin()+K, comps.end(), area-descending) -- the SET it leaves
// in the first K slots (cca.cpp:225-228), libstdc++ bits/stl_heap.h semantics.  The heap holds
// (area << 32 | component) words; only the area takes part in comparisons, like the reference's
// comparator (cca.cpp:179-185).  One CTA per image:
//   * all threads stream the components in ascending order in chunks, keeping only candidates
//     (area >= thres) that could still enter (area > current heap minimum, which never decreases);
//     survivors are compacted IN ORDER into a shared queue;
//   * thread 0 replays the queue sequentially through __pop_heap / __adjust_heap / __push_heap.
// ---------------------------------------------------------------------------------------------
#define SEL_CHUNK 4096  // components examined per round (4 per thread)

__device__ __forceinline__ uint32_t hs_area(unsigned long long e) { return (uint32_t)(e >> 32); }

__device__ void hs_push_heap(unsigned long long* h, int hole, int top, unsigned long long value) {
    int parent = (hole - 1) / 2;
    while (hole > top && hs_area(h[parent]) > hs_area(value)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}
__device__ void hs_adjust_heap(unsigned long long* h, int hole, int len, unsigned long long value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (hs_area(h[child]) > hs_area(h[child - 1])) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    hs_push_heap(h, hole, top, value);
}

// generic body shared by the pipeline kernel and the debug entry point
__device__ void heap_select_body(const uint32_t* __restrict__ area, int ncomp, int K, int thres,
                                 unsigned long long* heap, uint32_t* __restrict__ mark_out /* |= 1<<31 */,
                                 uint8_t* __restrict__ kept_bytes /* or nullptr */) {
    __shared__ unsigned long long s_queue[SEL_CHUNK];
    __shared__ int s_warp[32];
    __shared__ int s_qn, s_filled;
    __shared__ uint32_t s_min;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x;
    const int per = SEL_CHUNK / nt;  // consecutive components per thread
    if (tid == 0) {
        s_filled = 0;
        s_min = 0;
    }
    __syncthreads();
    for (int base = 0; base < ncomp; base += SEL_CHUNK) {
        const int filled = s_filled;
        const bool filling = filled < K;
        const uint32_t curmin = s_min;
        // thread t owns components base + t*per .. +per-1 (keeps ascending order inside the thread)
        unsigned long long mine[SEL_CHUNK / 256];
        int cnt = 0;
        for (int u = 0; u < per; u++) {
            const int c = base + tid * per + u;
            if (c < ncomp) {
                const uint32_t a = area[c];
                if ((int)a >= thres && (filling || a > curmin)) mine[cnt++] = ((unsigned long long)a << 32) | (uint32_t)c;
            }
        }
        // ordered compaction: exclusive scan of cnt over the CTA
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FSLIC_FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < (nt >> 5)) ? s_warp[lane] : 0;
            int z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(FSLIC_FULL, z, o);
                if (lane >= o) z += y;
            }
            s_warp[lane] = z - w;
            if (lane == 31) s_qn = z;
        }
        __syncthreads();
        const int pos = s_warp[warp] + x - cnt;
        for (int u = 0; u < cnt; u++) s_queue[pos + u] = mine[u];
        __syncthreads();
        const int qn = s_qn;
        // phase 1: the first K candidates fill the heap array in order
        int consumed = 0;
        if (filling) {
            consumed = min(qn, K - filled);
            for (int u = tid; u < consumed; u += nt) heap[filled + u] = s_queue[u];
            __syncthreads();
        }
        if (tid == 0) {
            int f = filled + consumed;
            if (filling && f == K) {  // __make_heap(first, middle)
                if (K >= 2) {
                    int parent = (K - 2) / 2;
                    for (;;) {
                        const unsigned long long value = heap[parent];
                        hs_adjust_heap(heap, parent, K, value);
                        if (parent == 0) break;
                        parent--;
                    }
                }
            }
            if (f == K) {
                // phase 2: the rest, one by one (cca.cpp:226 -> __heap_select loop)
                for (int u = consumed; u < qn; u++) {
                    const unsigned long long e = s_queue[u];
                    if (hs_area(e) > hs_area(heap[0])) hs_adjust_heap(heap, 0, K, e);  // __pop_heap(first, middle, i)
                }
                s_min = hs_area(heap[0]);
            }
            s_filled = f;
        }
        __syncthreads();
    }
    // publish the selected set
    const int filled = s_filled;
    for (int u = tid; u < filled; u += nt) {
        const uint32_t c = (uint32_t)(heap[u] & 0xffffffffu);
        if (mark_out) mark_out[c] |= 0x80000000u;
        if (kept_bytes) kept_bytes[c] = 1;
    }
}

__global__ void __launch_bounds__(1024) k_cca_select(CcaParams cp, uint32_t* __restrict__ carea_all,
                                                     CcaCounters* __restrict__ counters,
                                                     unsigned long long* __restrict__ heap_global) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    const int b = blockIdx.x;
    CcaCounters* ct = &counters[b];
    if (ct->ncand <= cp.K) return;  // cca.cpp:225: nothing to select
    unsigned long long* heap = cp.heap_in_smem ? reinterpret_cast<unsigned long long*>(sel_smem)
                                               : heap_global + (size_t)b * cp.K;
    heap_select_body(carea_all + (size_t)b * cp.N, ct->ncomp, cp.K, cp.thres, heap,
                     carea_all + (size_t)b * cp.N, nullptr);
    if (threadIdx.x == 0) ct->sel_mode = 1;
}

__global__ void __launch_bounds__(1024) k_debug_heap_select(const uint32_t* __restrict__ area, int n, int middle,
                                                            uint8_t* __restrict__ kept,
                                                            unsigned long long* __restrict__ heap_global) {
    heap_select_body(area, n, middle, 0, heap_global, nullptr, kept);
}

__device__ __forceinline__ bool cca_is_kept(uint32_t a, int sel_mode, int thres) {
    return sel_mode ? (a >> 31) : ((int)a >= thres);
}

__global__ void __launch_bounds__(CCA_BLOCK) k_kept_count(CcaParams cp, const uint32_t* __restrict__ carea_all,
                                                          const CcaCounters* __restrict__ counters,
                                                          int* __restrict__ blkcnt) {
    __shared__ int s_cnt;
    const int b = blockIdx.y;
    const int ncomp = counters[b].ncomp;
    if (blockIdx.x * CCA_BLOCK >= ncomp) return;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int c = blockIdx.x * CCA_BLOCK + threadIdx.x;
    const bool kept = (c < ncomp) && cca_is_kept(carea_all[(size_t)b * cp.N + c], counters[b].sel_mode, cp.thres);
    const unsigned m = __ballot_sync(FSLIC_FULL, kept);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0) blkcnt[(size_t)b * cp.nblk + blockIdx.x] = s_cnt;
}

// newlabel[c] = rank among kept (cca.cpp:234-236), 0xFFFF for components that must be absorbed
__global__ void __launch_bounds__(CCA_BLOCK) k_kept_label(CcaParams cp, const uint32_t* __restrict__ carea_all,
                                                          const CcaCounters* __restrict__ counters,
                                                          const int* __restrict__ blkoff,
                                                          uint16_t* __restrict__ cnew_all) {
    __shared__ int s_warp[32];
    const int b = blockIdx.y;
    const int ncomp = counters[b].ncomp;
    if (blockIdx.x * CCA_BLOCK >= ncomp) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x * CCA_BLOCK + tid;
    const bool kept = (c < ncomp) && cca_is_kept(carea_all[(size_t)b * cp.N + c], counters[b].sel_mode, cp.thres);
    const unsigned m = __ballot_sync(FSLIC_FULL, kept);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FSLIC_FULL, x, o);
            if (lane >= o) x += y;
        }
        s_warp[lane] = x - w;
    }
    __syncthreads();
    if (c < ncomp) {
        uint16_t v = 0xFFFF;
        if (kept) v = (uint16_t)(blkoff[(size_t)b * cp.nblk + blockIdx.x] + s_warp[warp] + __popc(m & ((1u << lane) - 1)));
        cnew_all[(size_t)b * cp.N + c] = v;
    }
}

// final label of every component, written at its leader pixel (cca.cpp:238-255)
__global__ void __launch_bounds__(256) k_cca_absorb(CcaParams cp, const int* __restrict__ par_all,
                                                    const uint32_t* __restrict__ aux_all,
                                                    const int* __restrict__ cleader_all,
                                                    const uint16_t* __restrict__ cnew_all,
                                                    const CcaCounters* __restrict__ counters,
                                                    uint16_t* __restrict__ final_all) {
    const int b = blockIdx.y;
    const int ncomp = counters[b].ncomp;
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c0 >= ncomp) return;
    const int* par = par_all + (size_t)b * cp.N;
    const uint32_t* aux = aux_all + (size_t)b * cp.N;
    const int* cleader = cleader_all + (size_t)b * cp.N;
    const uint16_t* cnew = cnew_all + (size_t)b * cp.N;
    int c = c0;
    uint16_t lab = cnew[c];
    while (lab == 0xFFFF) {
        if (c == 0) {  // subst[0] = 0 when component 0 was not kept (cca.cpp:238)
            lab = 0;
            break;
        }
        const int l = cleader[c];
        const int q = (l % cp.W > 0) ? (l - 1) : (l - cp.W);
        c = (int)aux[par[q]];  // component of the neighbour; strictly smaller than c
        lab = cnew[c];
    }
    final_all[(size_t)b * cp.N + cleader[c0]] = lab;
}

__global__ void __launch_bounds__(256) k_cca_output(CcaParams cp, const int* __restrict__ par_all,
                                                    const uint16_t* __restrict__ final_all,
                                                    uint16_t* __restrict__ out_all) {
    const int b = blockIdx.y;
    const int* par = par_all + (size_t)b * cp.N;
    const uint16_t* fin = final_all + (size_t)b * cp.N;
    uint16_t* out = out_all + (size_t)b * cp.N;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < cp.N; p += gridDim.x * blockDim.x) out[p] = fin[par[p]];
}
