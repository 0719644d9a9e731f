// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" door into the UNMODIFIED reference sources under
// /root/reference/src (compiled in place by oracle/Makefile, outputs into
// oracle/_ref/).  It drives the reference exactly the way cfast_slic.pyx does
// (cfast_slic.pyx:124-147 initialize, :150-197 iterate, :371-396
// enforce_connectivity) and additionally exposes the protected per-stage
// buffers (quad_image / assignment, context.h:48-50) so that every CUDA kernel
// has its own stage oracle.
#include <cstring>
#include <cstdint>
#include <string>
#include "context.h"
#include "cca.h"
#include "parallel.h"
#include "arch/x64/avx2.h"
#include "fast-slic.h"

namespace {
struct ProbeAvx2 : public fslic::Context_X64_AVX2 {
    ProbeAvx2(int H, int W, int K, const uint8_t* image, Cluster* clusters)
        : fslic::Context_X64_AVX2(H, W, K, image, clusters) {}
    void dump(uint8_t* quad_out, uint16_t* precca_out) {
        for (int i = 0; i < H; i++)
            for (int j = 0; j < W; j++) {
                if (quad_out)
                    for (int c = 0; c < 4; c++) quad_out[(i * W + j) * 4 + c] = quad_image.get(i, 4 * j + c);
                if (precca_out) precca_out[i * W + j] = assignment.get(i, j);
            }
    }
};
struct ProbeStd : public fslic::Context {
    ProbeStd(int H, int W, int K, const uint8_t* image, Cluster* clusters)
        : fslic::Context(H, W, K, image, clusters) {}
    void dump(uint8_t* quad_out, uint16_t* precca_out) {
        for (int i = 0; i < H; i++)
            for (int j = 0; j < W; j++) {
                if (quad_out)
                    for (int c = 0; c < 4; c++) quad_out[(i * W + j) * 4 + c] = quad_image.get(i, 4 * j + c);
                if (precca_out) precca_out[i * W + j] = assignment.get(i, j);
            }
    }
};
template <typename Ctx>
void configure(Ctx& ctx, float compactness, float min_size_factor, int stride, int convert_to_lab, int num_threads) {
    ctx.num_threads = num_threads;
    ctx.compactness = compactness;
    ctx.min_size_factor = min_size_factor;
    ctx.subsample_stride_config = (int16_t)stride;
    ctx.convert_to_lab = convert_to_lab != 0;
    ctx.preemptive = false;
    ctx.preemptive_thres = 0.05f;
    ctx.manhattan_spatial_dist = true;
    ctx.debug_mode = false;
}
// float-distance contexts (context.h:100-125), driven like cfast_slic.pyx:198-252: variant 0 = ContextRealDist
// ("standard"), 1 = ContextRealDistL2, 2 = ContextRealDistNoQ
template <typename Base>
struct ProbeReal : public Base {
    ProbeReal(int H, int W, int K, const uint8_t* image, Cluster* clusters) : Base(H, W, K, image, clusters) {}
    void dump(uint16_t* precca_out) {
        if (!precca_out) return;
        for (int i = 0; i < this->H; i++)
            for (int j = 0; j < this->W; j++) precca_out[i * this->W + j] = this->assignment.get(i, j);
    }
};
template <typename Ctx>
void run_real(Ctx& ctx, uint16_t* out, int max_iter, float compactness, float min_size_factor, int stride, int convert_to_lab,
              int num_threads, uint16_t* precca_out) {
    configure(ctx, compactness, min_size_factor, stride, convert_to_lab, num_threads);
    ctx.initialize_state();
    ctx.iterate(out, max_iter);
    ctx.dump(precca_out);
}
}  // namespace

extern "C" {

int ref_sizeof_cluster() { return (int)sizeof(Cluster); }

// cfast_slic.pyx:124-147
void ref_initialize(int H, int W, int K, const uint8_t* image, Cluster* clusters) {
    fslic::ContextBuilder builder("standard");
    fslic::Context* ctx = builder.build(H, W, K, image, clusters);
    ctx->initialize_clusters();
    delete ctx;
}

// cfast_slic.pyx:150-197 ; arch: 0 = "standard", 1 = "x64/avx2"
// quad_out (u8[H*W*4]) and precca_out (u16[H*W]) may be NULL.
void ref_iterate(int arch, int H, int W, int K, const uint8_t* image, Cluster* clusters, uint16_t* out,
                 int max_iter, float compactness, float min_size_factor, int stride, int convert_to_lab,
                 int num_threads, uint8_t* quad_out, uint16_t* precca_out) {
    if (arch == 1) {
        ProbeAvx2 ctx(H, W, K, image, clusters);
        configure(ctx, compactness, min_size_factor, stride, convert_to_lab, num_threads);
        ctx.initialize_state();
        ctx.iterate(out, max_iter);
        ctx.dump(quad_out, precca_out);
    } else {
        ProbeStd ctx(H, W, K, image, clusters);
        configure(ctx, compactness, min_size_factor, stride, convert_to_lab, num_threads);
        ctx.initialize_state();
        ctx.iterate(out, max_iter);
        ctx.dump(quad_out, precca_out);
    }
}

// the same call with preemptive = true (context.h:32-33; cfast_slic.pyx:183-184)
void ref_iterate_preemptive(int arch, int H, int W, int K, const uint8_t* image, Cluster* clusters, uint16_t* out,
                            int max_iter, float compactness, float min_size_factor, int stride, int convert_to_lab,
                            float preemptive_thres, int num_threads, uint16_t* precca_out) {
    if (arch == 1) {
        ProbeAvx2 ctx(H, W, K, image, clusters);
        configure(ctx, compactness, min_size_factor, stride, convert_to_lab, num_threads);
        ctx.preemptive = true;
        ctx.preemptive_thres = preemptive_thres;
        ctx.initialize_state();
        ctx.iterate(out, max_iter);
        ctx.dump(nullptr, precca_out);
    } else {
        ProbeStd ctx(H, W, K, image, clusters);
        configure(ctx, compactness, min_size_factor, stride, convert_to_lab, num_threads);
        ctx.preemptive = true;
        ctx.preemptive_thres = preemptive_thres;
        ctx.initialize_state();
        ctx.iterate(out, max_iter);
        ctx.dump(nullptr, precca_out);
    }
}

// cfast_slic.pyx:371-396 (K = max label + 1 is computed by the caller, as the pyx does)
void ref_enforce_connectivity(uint16_t* labels, int H, int W, int K, int min_threshold, int num_threads) {
    fsparallel::Scope scope(num_threads);
    cca::ConnectivityEnforcer ce(labels, H, W, K, min_threshold);
    ce.execute(labels);
}

// float-distance contexts, cfast_slic.pyx:198-252: variant 0 = ContextRealDist, 1 = ContextRealDistL2, 2 = ContextRealDistNoQ
void ref_iterate_real(int variant, int H, int W, int K, const uint8_t* image, Cluster* clusters, uint16_t* out, int max_iter,
                      float compactness, float min_size_factor, int stride, int convert_to_lab, int num_threads,
                      uint16_t* precca_out) {
    if (variant == 0) {
        ProbeReal<fslic::ContextRealDist> ctx(H, W, K, image, clusters);
        run_real(ctx, out, max_iter, compactness, min_size_factor, stride, convert_to_lab, num_threads, precca_out);
    } else if (variant == 1) {
        ProbeReal<fslic::ContextRealDistL2> ctx(H, W, K, image, clusters);
        run_real(ctx, out, max_iter, compactness, min_size_factor, stride, convert_to_lab, num_threads, precca_out);
    } else {
        ProbeReal<fslic::ContextRealDistNoQ> ctx(H, W, K, image, clusters);
        run_real(ctx, out, max_iter, compactness, min_size_factor, stride, convert_to_lab, num_threads, precca_out);
    }
}

// fast-slic.cpp:16-78 through the door cfast_slic.pyx:262-270 uses.  counts[K], neighbors[K * 12] (max_conn = 12).
void ref_get_connectivity(int H, int W, int K, const uint16_t* assignment, int32_t* counts, uint32_t* neighbors) {
    Connectivity* conn = fast_slic_get_connectivity(H, W, K, assignment);
    for (int k = 0; k < K; k++) {
        counts[k] = conn->num_neighbors[k];
        for (int t = 0; t < conn->num_neighbors[k] && t < 12; t++) neighbors[k * 12 + t] = conn->neighbors[k][t];
    }
    fast_slic_free_connectivity(conn);
}

// fast-slic.cpp:141-168 (cfast_slic.pyx:283-320)
void ref_get_mask_density(int H, int W, int K, const Cluster* clusters, const uint16_t* assignment, const uint8_t* mask,
                          uint8_t* densities) {
    fast_slic_get_mask_density(H, W, K, clusters, assignment, mask, densities);
}
void ref_cluster_density_to_mask(int H, int W, int K, const Cluster* clusters, const uint16_t* assignment,
                                 const uint8_t* densities, uint8_t* result) {
    fast_slic_cluster_density_to_mask(H, W, K, clusters, assignment, densities, result);
}

}  // extern "C"
