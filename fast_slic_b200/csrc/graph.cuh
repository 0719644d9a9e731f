// fast_slic_b200/csrc/graph.cuh -- consumers of the final label map.
//
// Replaces fast_slic_get_connectivity (/root/reference/src/fast-slic.cpp:16-78), fast_slic_get_mask_density (:141-155)
// and fast_slic_cluster_density_to_mask (:157-168).  fast_slic_knn_connectivity (:80-130) is not provided: it indexes
// its cell vector with a float expression that runs past the end for any centre low in the last cell row (:88), i.e.
// the reference itself has no defined result to match (oracle/slic_oracle.c).
#pragma once
#include "common.cuh"

#define CONN_MAX 12            // max_conn, fast-slic.cpp:17
#define CONN_EMPTY 0xffffffffu
#define CONN_NOORDER 0xffffffffffffffffull

// ---------------------------------------------------------------------------------------------
// Adjacency graph.  The reference scans the pixels in raster order; at every pixel it probes the right, lower and
// lower-right neighbour and links the two labels unless they are linked already or either list is full (12).  A
// refused link is refused again at every later probe (lists only grow), so the outcome is a function of the FIRST
// probe of every unordered label pair, taken in scan order:
//   k_conn_discover  every probe whose labels differ files  min(order)  under its pair in an open-addressing hash
//                    table (order = 3 * pixel + probe slot; CAS claims a slot, atomicMin keeps the first probe);
//   (radix sort)     the table sorted by order = the distinct pairs in the order the reference first meets them;
//   k_conn_walk      one thread replays that short list (a few pairs per superpixel) with the capacity rule.
// If the table overflows (label maps with millions of distinct adjacent pairs) k_conn_scan replays the reference's
// loop itself, one thread over all pixels -- slow, exact, and never needed for superpixel maps.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t conn_hash(uint32_t key) {
    key ^= key >> 15;
    key *= 0x2c1b3c6du;
    key ^= key >> 12;
    key *= 0x297a2d39u;
    key ^= key >> 15;
    return key;
}

__global__ void __launch_bounds__(256) k_conn_discover(const uint16_t* __restrict__ lab, int H, int W, int K,
                                                        uint32_t* __restrict__ tkey, unsigned long long* __restrict__ tord,
                                                        uint32_t tmask, int* __restrict__ overflow) {
    const long n = (long)(H - 1) * (W - 1);
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t / (W - 1)), j = (int)(t - (long)i * (W - 1));
        const long p = (long)i * W + j;
        const uint32_t s = lab[p];
        if (s >= (uint32_t)K) continue;
        const uint32_t nb[3] = {lab[p + 1], lab[p + W], lab[p + W + 1]};
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const uint32_t g = nb[u];
            if (g >= (uint32_t)K || g == s) continue;
            if (u == 2 && (g == nb[0] || g == nb[1])) continue;  // same pair, later order: cannot be the first probe
            if (u == 1 && g == nb[0]) continue;
            const uint32_t key = s < g ? (s << 16 | g) : (g << 16 | s);
            const unsigned long long order = (unsigned long long)p * 3ull + (unsigned long long)u;
            uint32_t h = conn_hash(key) & tmask;
            int probes = 0;
            for (;;) {
                const uint32_t old = atomicCAS(&tkey[h], CONN_EMPTY, key);
                if (old == CONN_EMPTY || old == key) {
                    atomicMin(&tord[h], order);
                    break;
                }
                h = (h + 1) & tmask;
                if (++probes > 512) {
                    *overflow = 1;
                    break;
                }
            }
        }
    }
}

// sorted_ord / sorted_key: the table sorted by order; the first CONN_NOORDER entry ends the list
__global__ void k_conn_walk(const unsigned long long* __restrict__ sorted_ord, const uint32_t* __restrict__ sorted_key,
                            uint32_t tsize, int K, int32_t* __restrict__ counts, uint32_t* __restrict__ neighbors) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint32_t e = 0; e < tsize; e++) {
        if (sorted_ord[e] == CONN_NOORDER) break;
        const uint32_t key = sorted_key[e];
        const uint32_t a = key >> 16, b = key & 0xffffu;
        const int na = counts[a], nb = counts[b];
        if (na >= CONN_MAX || nb >= CONN_MAX) continue;  // fast-slic.cpp:43
        neighbors[a * CONN_MAX + na] = b;
        neighbors[b * CONN_MAX + nb] = a;
        counts[a] = na + 1;
        counts[b] = nb + 1;
    }
}

// exact replay of fast-slic.cpp:29-74 by one thread (table overflow only)
__global__ void k_conn_scan(const uint16_t* __restrict__ lab, int H, int W, int K, int32_t* __restrict__ counts,
                            uint32_t* __restrict__ neighbors) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int i = 0; i < H - 1; i++) {
        for (int j = 0; j < W - 1; j++) {
            const long p = (long)i * W + j;
            const uint32_t s = lab[p];
            if (s >= (uint32_t)K) continue;
            int ns = counts[s];
            const long probe[3] = {p + 1, p + W, p + W + 1};
            for (int u = 0; u < 3; u++) {
                const uint32_t g = lab[probe[u]];
                if (g >= (uint32_t)K || g == s) continue;
                const int ng = counts[g];
                if (ns >= CONN_MAX || ng >= CONN_MAX) continue;
                bool exists = false;
                for (int v = 0; v < ns && !exists; v++) exists = neighbors[s * CONN_MAX + v] == g;
                for (int v = 0; v < ng && !exists; v++) exists = neighbors[g * CONN_MAX + v] == s;
                if (exists) continue;
                neighbors[g * CONN_MAX + ng] = s;
                counts[g] = ng + 1;
                neighbors[s * CONN_MAX + ns] = g;
                ns++;
            }
            counts[s] = ns;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Mask density (fast-slic.cpp:141-155): sum[k] = sum of mask over the pixels labelled k; density = min(255, sum /
// max(num_members, 1)) -- num_members being the Cluster field (the last subsampled update's count), as in the
// reference.  Lanes holding the same label add once (MATCH.ANY + REDUX).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mask_sum(const uint16_t* __restrict__ lab, const uint8_t* __restrict__ mask, long n,
                                                   int K, int32_t* __restrict__ sum) {
    const long step = (long)gridDim.x * blockDim.x;
    const long nround = (n + step - 1) / step * step;  // whole warps stay in the loop: the warp intrinsics need all lanes
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += step) {
        uint32_t l = 0xffffffffu, v = 0;
        if (t < n) {
            l = lab[t];
            v = mask[t];
            if (l >= (uint32_t)K) l = 0xffffffffu;
        }
        const unsigned peers = __match_any_sync(FSLIC_FULL, l);
        const uint32_t total = __reduce_add_sync(peers, v);
        if (l != 0xffffffffu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1) && total) atomicAdd(&sum[l], (int)total);
    }
}

__global__ void k_density_final(const int32_t* __restrict__ sum, const fslic_cluster* __restrict__ clusters, int K,
                                uint8_t* __restrict__ dens) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const uint32_t den = clusters[k].num_members > 1u ? clusters[k].num_members : 1u;
    const uint32_t v = (uint32_t)sum[k] / den;
    dens[k] = (uint8_t)(v < 255u ? v : 255u);
}

// fast-slic.cpp:157-168
__global__ void __launch_bounds__(256) k_density_broadcast(const uint16_t* __restrict__ lab, const uint8_t* __restrict__ dens,
                                                            long n, int K, uint8_t* __restrict__ out) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const uint32_t l = lab[t];
        out[t] = l < (uint32_t)K ? dens[l] : (uint8_t)0;
    }
}
