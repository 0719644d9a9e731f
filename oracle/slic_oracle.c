/*
 * oracle/slic_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's default
 * `Slic.iterate()` hot path (Algy/fast-slic @ e6f6b4f).  It exists so that the
 * CUDA path can be checked bit-for-bit on machines where /root/reference does
 * not exist (the GPU box).  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and only as the checker.
 *
 * Parity pinning: this restatement is itself checked against the compiled,
 * unmodified reference (oracle/_ref/libfslic_ref.so, built by oracle/Makefile)
 * in tests/test_oracle_vs_ref.py and against the committed golden vectors in
 * tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float y, x, r, g, b, a;
    uint16_t number;
    uint8_t is_active, is_updatable;
    uint32_t num_members;
} OrcCluster; /* src/fast-slic-common.h:10-23 -- 32 bytes */

int orc_sizeof_cluster(void) { return (int)sizeof(OrcCluster); }

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int ceil_int(int a, int b) { return (a + b - 1) / b; }   /* fast-slic-common.h:58-60 */
static int round_int(int a, int b) { return (a + b / 2) / b; }  /* fast-slic-common.h:63-65 */

/* ------------------------------------------------------------------ */
/* RGB -> CIELAB integer LUT pipeline: src/cielab.h:286-325            */
/* ------------------------------------------------------------------ */
#define SRGB_SHIFT 13
#define SRGB_MAX (1 << SRGB_SHIFT)
#define LAB_SHIFT 16
#define OUTPUT_SHIFT 1

/* cielab.h:12-20 documents the generating formula of _srgb_gamma_tbl (cielab.h:22-279);
 * evaluating it in double reproduces every literal of that table exactly (checked in
 * tests/test_oracle_vs_ref.py through the full 2^24 colour cube). */
static float srgb_gamma(int a) {
    double v = a / 255.0;
    double X = (v <= 0.04045) ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4);
    return (float)X;
}

static float lab_nonlin(float v) { /* cielab.h:328-332 */
    float lo = 7.787f * v + 0.137931f;
    float hi = powf(v, 0.333333f);
    return (v > 0.008856f) ? hi : lo;
}

/* gamma[256], labtbl[8193], Cb[9] : cielab.h:297-305 */
void orc_lab_tables(int32_t* gamma, int32_t* labtbl, int32_t* Cb) {
    static const float C[9] = {0.43395633f, 0.37621531f, 0.18984309f, 0.2126729f, 0.7151522f,
                               0.072175f,   0.01775782f, 0.1094756f,  0.87283638f}; /* cielab.h:288-292 */
    for (int i = 0; i < 256; i++) gamma[i] = (int)(srgb_gamma(i) * SRGB_MAX);
    for (int i = 0; i < 9; i++) Cb[i] = (int)roundf(C[i] * (1 << LAB_SHIFT));
    for (int i = 0; i <= SRGB_MAX; i++) labtbl[i] = (int32_t)roundf(lab_nonlin((float)i / SRGB_MAX) * SRGB_MAX);
}

static int32_t g_gamma[256], g_labtbl[SRGB_MAX + 1], g_Cb[9];
static int g_tables_ready = 0;
static void ensure_tables(void) {
    if (!g_tables_ready) {
        orc_lab_tables(g_gamma, g_labtbl, g_Cb);
        g_tables_ready = 1;
    }
}

/* cielab.h:308-325 */
static void lab_convert(uint8_t R, uint8_t G, uint8_t B, uint8_t* l, uint8_t* a, uint8_t* b) {
    int sr = g_gamma[R], sg = g_gamma[G], sb = g_gamma[B];
    int xr = (g_Cb[0] * sr + g_Cb[1] * sg + g_Cb[2] * sb) >> LAB_SHIFT;
    int yr = (g_Cb[3] * sr + g_Cb[4] * sg + g_Cb[5] * sb) >> LAB_SHIFT;
    int zr = (g_Cb[6] * sr + g_Cb[7] * sg + g_Cb[8] * sb) >> LAB_SHIFT;
    int fx = g_labtbl[xr], fy = g_labtbl[yr], fz = g_labtbl[zr];
    int ciel = 116 * fy - (16 << SRGB_SHIFT);
    int ciea = 500 * (fx - fy) + (128 << SRGB_SHIFT);
    int cieb = 200 * (fy - fz) + (128 << SRGB_SHIFT);
    /* the reference shifts as unsigned, subtracts as unsigned, then clamps as int */
    *l = (uint8_t)clampi((int)((unsigned)ciel >> (SRGB_SHIFT - OUTPUT_SHIFT)), 0, 255);
    *a = (uint8_t)clampi((int)(((unsigned)ciea >> (SRGB_SHIFT - OUTPUT_SHIFT)) - (64u << OUTPUT_SHIFT)), 0, 255);
    *b = (uint8_t)clampi((int)(((unsigned)cieb >> (SRGB_SHIFT - OUTPUT_SHIFT)) - (64u << OUTPUT_SHIFT)), 0, 255);
}

/* cielab.h:337-353 (convert_to_lab) / context.cpp:118-127 (raw RGB); quad = u8[H*W*4], alpha 0 */
void orc_rgb_to_quad(const uint8_t* rgb, int H, int W, int convert_to_lab, uint8_t* quad) {
    ensure_tables();
    for (long p = 0; p < (long)H * W; p++) {
        if (convert_to_lab) {
            lab_convert(rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2], &quad[4 * p], &quad[4 * p + 1], &quad[4 * p + 2]);
        } else {
            quad[4 * p] = rgb[3 * p];
            quad[4 * p + 1] = rgb[3 * p + 1];
            quad[4 * p + 2] = rgb[3 * p + 2];
        }
        quad[4 * p + 3] = 0;
    }
}

/* ------------------------------------------------------------------ */
/* grid seeding: src/context.cpp:43-97                                 */
/* ------------------------------------------------------------------ */
void orc_initialize_clusters(int H, int W, int K, const uint8_t* image, OrcCluster* clusters) {
    if (H <= 0 || W <= 0 || K <= 0) return;
    int n_y = (int)sqrt((double)K);
    int* n_xs = (int*)malloc(sizeof(int) * n_y);
    for (int i = 0; i < n_y; i++) n_xs[i] = K / n_y;
    int remainder = K % n_y, row = 0;
    while (remainder-- > 0) {
        n_xs[row]++;
        row += 2;
        if (row >= n_y) row = 1 % n_y;
    }
    int h = ceil_int(H, n_y), acc_k = 0;
    for (int i = 0; i < H; i += h) {
        int bi = i / h;
        if (bi > n_y - 1) bi = n_y - 1;
        int w = ceil_int(W, n_xs[bi]);
        for (int j = 0; j < W; j += w) {
            if (acc_k >= K) break;
            clusters[acc_k].y = (float)clampi(i + h / 2, 0, H - 1);
            clusters[acc_k].x = (float)clampi(j + w / 2, 0, W - 1);
            clusters[acc_k].is_active = 1;
            clusters[acc_k].is_updatable = 1;
            acc_k++;
        }
    }
    while (acc_k < K) {
        clusters[acc_k].is_active = 1;
        clusters[acc_k].is_updatable = 1;
        clusters[acc_k].y = (float)(H / 2);
        clusters[acc_k].x = (float)(W / 2);
        acc_k++;
    }
    for (int k = 0; k < K; k++) {
        /* context.cpp:88 computes `W * clusters[k].y + clusters[k].x` on FLOAT operands; with the reference's own
         * build flags (setup.py:137-149: -mfma, GCC's default -ffp-contract=fast) that is ONE fused multiply-add,
         * truncated to int.  Above 2^24 pixels the float result is not the exact pixel index any more (odd indices
         * round to a neighbour): reproduced here with fmaf(), which is exactly that instruction's arithmetic. */
        int base = (int)fmaf((float)W, clusters[k].y, clusters[k].x);
        clusters[k].r = image[3 * base];
        clusters[k].g = image[3 * base + 1];
        clusters[k].b = image[3 * base + 2];
        clusters[k].number = (uint16_t)k;
        clusters[k].num_members = 0;
    }
    free(n_xs);
}

/* ------------------------------------------------------------------ */
/* spatial term: src/context.cpp:23-40 (manhattan branch). The patch   */
/* value depends only on m = |di| + |dj|, so a 1-D table of 2S+1.      */
/* ------------------------------------------------------------------ */
void orc_spatial_lut(int S, float compactness, int color_shift, uint16_t* lut) {
    float coef = 1.0f / ((float)S / compactness);
    coef *= (float)(1 << color_shift);
    for (int m = 0; m <= 2 * S; m++) lut[m] = (uint16_t)(coef * (float)m);
}

/* ------------------------------------------------------------------ */
/* assign: src/context.cpp:200-243 (scheduler) + :259-298 (kernel).    */
/* Restated serially in the reference's own scatter form: phases 0..3, */
/* cells of that phase row-major, clusters of a cell in ascending k,   */
/* strict '<' against min_dists.  Windows are clipped to the image,    */
/* which is what the reference's zero padding amounts to (padding is   */
/* never read back).                                                   */
/* ------------------------------------------------------------------ */
static void assign_pass(int H, int W, int K, int S, OrcCluster* clusters, const uint8_t* quad, const uint16_t* lut,
                        uint16_t* assignment, uint16_t* min_dists, int stride, int rem) {
    for (long p = 0; p < (long)H * W; p++) min_dists[p] = 0xFFFF; /* :201-206 */
    for (int k = 0; k < K; k++) {                                 /* :209-212 */
        float x = clusters[k].x, y = clusters[k].y;
        clusters[k].x = x < 0 ? 0 : (x > (float)(W - 1) ? (float)(W - 1) : x);
        clusters[k].y = y < 0 ? 0 : (y > (float)(H - 1) ? (float)(H - 1) : y);
    }
    int T = 2 * S + 32;
    int cell_W = ceil_int(W, T), cell_H = ceil_int(H, T);
    for (int phase = 0; phase < 4; phase++) {
        for (int ci = phase / 2; ci < cell_H; ci += 2) {
            for (int cj = phase % 2; cj < cell_W; cj += 2) {
                for (int k = 0; k < K; k++) {
                    if (!clusters[k].is_active) continue;
                    int y = (int)clusters[k].y, x = (int)clusters[k].x;
                    if (y / T != ci || x / T != cj) continue;
                    int16_t cy = (int16_t)clusters[k].y, cx = (int16_t)clusters[k].x;
                    int16_t cr = (int16_t)clusters[k].r, cg = (int16_t)clusters[k].g, cb = (int16_t)clusters[k].b;
                    for (int i = cy - S; i <= cy + S; i++) {
                        if (i < 0 || i >= H) continue;
                        if (i % stride != rem) continue;
                        for (int j = cx - S; j <= cx + S; j++) {
                            if (j < 0 || j >= W) continue;
                            long p = (long)i * W + j;
                            int r = quad[4 * p], g = quad[4 * p + 1], b = quad[4 * p + 2];
                            uint16_t d = (uint16_t)(abs(r - cr) + abs(g - cg) + abs(b - cb) +
                                                    lut[abs(i - cy) + abs(j - cx)]);
                            if (min_dists[p] > d) {
                                min_dists[p] = d;
                                assignment[p] = clusters[k].number;
                            }
                        }
                    }
                }
            }
        }
    }
}

/* A faster equivalent of assign_pass used for large images: buckets the clusters per cell first
 * (context.cpp:214-221) instead of rescanning all K for every cell.  Same visiting order. */
static void assign_pass_bucketed(int H, int W, int K, int S, OrcCluster* clusters, const uint8_t* quad,
                                 const uint16_t* lut, uint16_t* assignment, uint16_t* min_dists, int stride,
                                 int rem) {
    for (long p = 0; p < (long)H * W; p++) min_dists[p] = 0xFFFF;
    for (int k = 0; k < K; k++) {
        float x = clusters[k].x, y = clusters[k].y;
        clusters[k].x = x < 0 ? 0 : (x > (float)(W - 1) ? (float)(W - 1) : x);
        clusters[k].y = y < 0 ? 0 : (y > (float)(H - 1) ? (float)(H - 1) : y);
    }
    int T = 2 * S + 32;
    int cell_W = ceil_int(W, T), cell_H = ceil_int(H, T);
    int ncell = cell_W * cell_H;
    int* start = (int*)calloc((size_t)ncell + 1, sizeof(int));
    int* items = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
    for (int k = 0; k < K; k++) {
        if (!clusters[k].is_active) continue;
        start[cell_W * ((int)clusters[k].y / T) + ((int)clusters[k].x / T) + 1]++;
    }
    for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
    int* fill = (int*)malloc(sizeof(int) * (size_t)(ncell > 0 ? ncell : 1));
    memcpy(fill, start, sizeof(int) * (size_t)ncell);
    for (int k = 0; k < K; k++) { /* ascending k inside each cell */
        if (!clusters[k].is_active) continue;
        items[fill[cell_W * ((int)clusters[k].y / T) + ((int)clusters[k].x / T)]++] = k;
    }
    for (int phase = 0; phase < 4; phase++)
        for (int ci = phase / 2; ci < cell_H; ci += 2)
            for (int cj = phase % 2; cj < cell_W; cj += 2) {
                int cell = ci * cell_W + cj;
                for (int t = start[cell]; t < start[cell + 1]; t++) {
                    int k = items[t];
                    int16_t cy = (int16_t)clusters[k].y, cx = (int16_t)clusters[k].x;
                    int16_t cr = (int16_t)clusters[k].r, cg = (int16_t)clusters[k].g, cb = (int16_t)clusters[k].b;
                    int i0 = cy - S < 0 ? 0 : cy - S, i1 = cy + S >= H ? H - 1 : cy + S;
                    int j0 = cx - S < 0 ? 0 : cx - S, j1 = cx + S >= W ? W - 1 : cx + S;
                    for (int i = i0; i <= i1; i++) {
                        if (i % stride != rem) continue;
                        for (int j = j0; j <= j1; j++) {
                            long p = (long)i * W + j;
                            int r = quad[4 * p], g = quad[4 * p + 1], b = quad[4 * p + 2];
                            uint16_t d = (uint16_t)(abs(r - cr) + abs(g - cg) + abs(b - cb) +
                                                    lut[abs(i - cy) + abs(j - cx)]);
                            if (min_dists[p] > d) {
                                min_dists[p] = d;
                                assignment[p] = clusters[k].number;
                            }
                        }
                    }
                }
            }
    free(start);
    free(items);
    free(fill);
}

/* ------------------------------------------------------------------ */
/* update: src/context.cpp:302-387 (all-active, quantised branch)      */
/* ------------------------------------------------------------------ */
/* active_grid == NULL: every pixel counts (preemptive off, or preemptive_grid.all_active(), context.cpp:314-328);
 * otherwise only pixels whose 2S x 2S cell is active (context.cpp:329-344, preemptive.h:109-111). */
static void update_pass_masked(int H, int W, int K, OrcCluster* clusters, const uint8_t* quad, const uint16_t* assignment,
                               int stride, int rem, const int* active_grid, int cell_pitch, int CW) {
    int32_t* n = (int32_t*)calloc((size_t)K, sizeof(int32_t));
    int32_t* acc = (int32_t*)calloc((size_t)K * 5, sizeof(int32_t));
    for (int i = rem; i < H; i += stride) { /* fit_to_stride(0) == rem, context.h:78-82 */
        for (int j = 0; j < W; j++) {
            long p = (long)i * W + j;
            if (active_grid && !active_grid[CW * (i / cell_pitch) + (j / cell_pitch)]) continue;
            uint16_t c = assignment[p];
            if (c == 0xFFFF) continue;
            n[c]++;
            acc[5 * c + 0] += i;
            acc[5 * c + 1] += j;
            acc[5 * c + 2] += quad[4 * p];
            acc[5 * c + 3] += quad[4 * p + 1];
            acc[5 * c + 4] += quad[4 * p + 2];
        }
    }
    for (int k = 0; k < K; k++) {
        OrcCluster* c = &clusters[k];
        if (!c->is_updatable) continue;
        c->num_members = (uint32_t)n[k];
        if (n[k] == 0) continue;
        c->y = (float)round_int(acc[5 * k + 0], n[k]);
        c->x = (float)round_int(acc[5 * k + 1], n[k]);
        c->r = (float)round_int(acc[5 * k + 2], n[k]);
        c->g = (float)round_int(acc[5 * k + 3], n[k]);
        c->b = (float)round_int(acc[5 * k + 4], n[k]);
    }
    free(n);
    free(acc);
}
static void update_pass(int H, int W, int K, OrcCluster* clusters, const uint8_t* quad, const uint16_t* assignment,
                        int stride, int rem) {
    update_pass_masked(H, W, K, clusters, quad, assignment, stride, rem, NULL, 1, 0);
}

/* ------------------------------------------------------------------ */
/* PreemptiveGrid::set_new_clusters, src/preemptive.h:113-177.          */
/* old_yx: the centres before this update (set_old_clusters, :105-108). */
/* Cells are 2S x 2S pixels (:37-56).  A cluster that moved less than   */
/* l1_thres counts its is_updatable down (2 -> 1 -> 0, and 0 stays 0);  */
/* every cluster within a 2S Chebyshev distance (truncated centres) of  */
/* a still-updatable one is active, and so is its cell.  Returns        */
/* b_all_active.  The L1 movement is a float |.| here; the centres are  */
/* integer valued in the quantised contexts, so whether the reference's */
/* unqualified abs() resolves to the int or the float overload cannot   */
/* change the result.                                                   */
/* ------------------------------------------------------------------ */
static int set_new_clusters(int H, int W, int K, int S, float thres, OrcCluster* clusters, const float* old_yx,
                            int* active_grid) {
    int CW = ceil_int(W, 2 * S), CH = ceil_int(H, 2 * S);
    for (int c = 0; c < CW * CH; c++) active_grid[c] = 0;
    for (int k = 0; k < K; k++) clusters[k].is_active = 0;                              /* :119-122 */
    float l1_thres = roundf((float)(2 * S) * thres);                                    /* :126 */
    if (l1_thres < 1.0f) l1_thres = 1.0f;
    for (int k = 0; k < K; k++) {                                                       /* :132-141 */
        if (!clusters[k].is_updatable) continue;
        float l1 = fabsf(old_yx[2 * k + 1] - clusters[k].x) + fabsf(old_yx[2 * k] - clusters[k].y);
        if (l1 < l1_thres) clusters[k].is_updatable--;
        else clusters[k].is_updatable = 2;
    }
    for (int k = 0; k < K; k++) {                                                       /* :143-168 */
        if (!clusters[k].is_updatable) continue;
        int y = (int)clusters[k].y, x = (int)clusters[k].x;
        int cy = y / (2 * S), cx = x / (2 * S);
        for (int n = 0; n < K; n++) {  /* the reference walks the 3 x 3 cells around (cy, cx); same set, see below */
            int ny = (int)clusters[n].y, nx = (int)clusters[n].x;
            int ncy = ny / (2 * S), ncx = nx / (2 * S);
            if (abs(ncy - cy) > 1 || abs(ncx - cx) > 1) continue;                       /* :150-156 */
            if (abs(ny - y) <= 2 * S && abs(nx - x) <= 2 * S) {                         /* :160-161 */
                clusters[n].is_active = 1;
                active_grid[CW * ncy + ncx] = 1;
            }
        }
    }
    (void)CH;
    int num_active = 0;
    for (int k = 0; k < K; k++) num_active += clusters[k].is_active;                    /* :170-176 */
    return num_active == K;
}

/* ------------------------------------------------------------------ */
/* libstdc++ std::partial_sort set semantics (bits/stl_heap.h,         */
/* bits/stl_algo.h __heap_select) restated for comparator              */
/* comp(a,b) := area[a] > area[b]  (src/cca.cpp:179-185, :225-228).    */
/* Only the SET held by the first `middle` slots matters afterwards    */
/* (cca.cpp:229 sorts them by leader).  Checked against the real       */
/* std::partial_sort in oracle/stl_probe.cpp by tests/test_heap_select */
/* ------------------------------------------------------------------ */
static void hs_push_heap(int* first, long hole, long top, int value, const int* area) {
    long parent = (hole - 1) / 2;
    while (hole > top && area[first[parent]] > area[value]) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void hs_adjust_heap(int* first, long hole, long len, int value, const int* area) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (area[first[child]] > area[first[child - 1]]) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    hs_push_heap(first, hole, top, value, area);
}
/* comps[0..n) -> after the call comps[0..middle) holds the selected set (heap order). */
void orc_heap_select(int* comps, long n, long middle, const int* area) {
    if (middle >= 2) { /* __make_heap */
        long parent = (middle - 2) / 2;
        for (;;) {
            int value = comps[parent];
            hs_adjust_heap(comps, parent, middle, value, area);
            if (parent == 0) break;
            parent--;
        }
    }
    for (long i = middle; i < n; i++) {
        if (area[comps[i]] > area[comps[0]]) { /* comp(i, first) */
            int value = comps[i];              /* __pop_heap(first, middle, i) */
            comps[i] = comps[0];
            hs_adjust_heap(comps, 0, middle, value, area);
        }
    }
}

/* ------------------------------------------------------------------ */
/* connectivity enforcement: src/cca.cpp:178-265                       */
/* (assign_disjoint_set :33-101, DisjointSet::merge cca.h:36-57,       */
/*  flatten :103-173).  Components are 4-connected regions of equal    */
/*  u16 value; the root of each is its minimum raster index.           */
/* ------------------------------------------------------------------ */
static int uf_find(int* parent, int x) {
    int r = x;
    while (parent[r] != r) r = parent[r];
    while (parent[x] != r) {
        int nx = parent[x];
        parent[x] = r;
        x = nx;
    }
    return r;
}
static void uf_union(int* parent, int a, int b) {
    int ra = uf_find(parent, a), rb = uf_find(parent, b);
    if (ra == rb) return;
    if (ra < rb) parent[rb] = ra; /* keep the minimum raster index as root (cca.h:38-55) */
    else parent[ra] = rb;
}
static int cmp_int(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    return (x > y) - (x < y);
}

void orc_enforce_connectivity(uint16_t* labels, int H, int W, int K, int min_threshold) {
    if (K <= 0 || H <= 0 || W <= 0) return; /* context.cpp:17 */
    long N = (long)H * W;
    int* parent = (int*)malloc(sizeof(int) * (size_t)N);
    for (long p = 0; p < N; p++) parent[p] = (int)p;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++) {
            int p = i * W + j;
            if (j > 0 && labels[p - 1] == labels[p]) uf_union(parent, p - 1, p);
            if (i > 0 && labels[p - W] == labels[p]) uf_union(parent, p - W, p);
        }
    /* flatten: component number = rank of root in raster order (cca.cpp:118-134) */
    int* comp = (int*)malloc(sizeof(int) * (size_t)N);
    int ncomp = 0;
    for (long p = 0; p < N; p++)
        if (parent[p] == p) comp[p] = ncomp++;
    int* area = (int*)calloc((size_t)ncomp, sizeof(int));
    int* leader = (int*)malloc(sizeof(int) * (size_t)ncomp);
    for (long p = 0; p < N; p++) {
        int r = uf_find(parent, (int)p);
        comp[p] = comp[r];
        area[comp[p]]++;
        if (r == p) leader[comp[p]] = (int)p;
    }
    uint16_t* subst = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)ncomp);
    for (int c = 0; c < ncomp; c++) subst[c] = 0xFFFF;
    int* cands = (int*)malloc(sizeof(int) * (size_t)ncomp);
    long ncand = 0;
    for (int c = 0; c < ncomp; c++) /* cca.cpp:213-219 */
        if (area[c] >= min_threshold) cands[ncand++] = c;
    if ((long)K < ncand) { /* cca.cpp:225-228 */
        orc_heap_select(cands, ncand, K, area);
        ncand = K;
    }
    qsort(cands, (size_t)ncand, sizeof(int), cmp_int); /* by leader == by component number (cca.cpp:229) */
    for (long r = 0; r < ncand; r++) subst[cands[r]] = (uint16_t)r; /* :234-236 */
    if (ncomp > 0 && subst[0] == 0xFFFF) subst[0] = 0;              /* :238 */
    for (int c = 0; c < ncomp; c++) {                               /* :240-255 */
        if (subst[c] != 0xFFFF) continue;
        int l = leader[c];
        uint16_t s = (l % W > 0) ? subst[comp[l - 1]] : subst[comp[l - W]];
        if (s == 0xFFFF) s = 0;
        subst[c] = s;
    }
    for (long p = 0; p < N; p++) labels[p] = subst[comp[p]]; /* :260-263 */
    free(parent); free(comp); free(area); free(leader); free(subst); free(cands);
}

/* ------------------------------------------------------------------ */
/* the whole pipeline: src/context.cpp:109-197                         */
/* quad_out (u8[H*W*4]) / precca_out (u16[H*W]) may be NULL.           */
/* ------------------------------------------------------------------ */
void orc_iterate_preemptive(int H, int W, int K, const uint8_t* image, OrcCluster* clusters, uint16_t* out, int max_iter,
                            float compactness, float min_size_factor, int stride, int convert_to_lab, int preemptive,
                            float preemptive_thres, uint8_t* quad_out, uint16_t* precca_out);
void orc_iterate(int H, int W, int K, const uint8_t* image, OrcCluster* clusters, uint16_t* out, int max_iter,
                 float compactness, float min_size_factor, int stride, int convert_to_lab, uint8_t* quad_out,
                 uint16_t* precca_out) {
    orc_iterate_preemptive(H, W, K, image, clusters, out, max_iter, compactness, min_size_factor, stride, convert_to_lab, 0,
                           0.05f, quad_out, precca_out);
}

/* preemptive != 0: BaseContext with preemptive = true (context.h:32-33, preemptive.h) */
void orc_iterate_preemptive(int H, int W, int K, const uint8_t* image, OrcCluster* clusters, uint16_t* out, int max_iter,
                            float compactness, float min_size_factor, int stride, int convert_to_lab, int preemptive,
                            float preemptive_thres, uint8_t* quad_out, uint16_t* precca_out) {
    if (H <= 0 || W <= 0 || K <= 0) return;
    int S = (int16_t)sqrt((double)(H * W / K)); /* context.h:60 -- integer division first */
    long N = (long)H * W;
    uint8_t* quad = (uint8_t*)malloc((size_t)N * 4);
    uint16_t* assignment = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)N);
    uint16_t* min_dists = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)N);
    uint16_t* lut = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(2 * S + 1));
    int color_shift = convert_to_lab ? OUTPUT_SHIFT : 0;
    orc_rgb_to_quad(image, H, W, convert_to_lab, quad); /* :114-127 */
    for (int k = 0; k < K; k++) {                       /* :128-135 */
        int y = clampi((int)clusters[k].y, 0, H - 1), x = clampi((int)clusters[k].x, 0, W - 1);
        clusters[k].r = quad[4 * ((long)y * W + x)];
        clusters[k].g = quad[4 * ((long)y * W + x) + 1];
        clusters[k].b = quad[4 * ((long)y * W + x) + 2];
    }
    for (long p = 0; p < N; p++) assignment[p] = 0xFFFF; /* :139-146 */
    orc_spatial_lut(S, compactness, color_shift, lut);    /* :147 */
    for (int k = 0; k < K; k++) clusters[k].is_updatable = 2; /* preemptive.h:59-67 (cooldown) */
    int all_active = 1;                                        /* preemptive.h:63 */
    int* active_grid = NULL;
    float* old_yx = NULL;
    if (preemptive && S > 0) {
        active_grid = (int*)calloc((size_t)ceil_int(W, 2 * S) * ceil_int(H, 2 * S), sizeof(int));
        old_yx = (float*)malloc(sizeof(float) * 2 * (size_t)K);
    }
    int rem = 0;
    for (int it = 0; it < max_iter; it++) { /* :158-175 */
        assign_pass_bucketed(H, W, K, S, clusters, quad, lut, assignment, min_dists, stride, rem);
        if (active_grid) {
            for (int k = 0; k < K; k++) { /* set_old_clusters, context.cpp:303 (after assign()'s clamp) */
                old_yx[2 * k] = clusters[k].y;
                old_yx[2 * k + 1] = clusters[k].x;
            }
            update_pass_masked(H, W, K, clusters, quad, assignment, stride, rem, all_active ? NULL : active_grid, 2 * S,
                               ceil_int(W, 2 * S));
            all_active = set_new_clusters(H, W, K, S, preemptive_thres, clusters, old_yx, active_grid); /* :385 */
        } else {
            update_pass(H, W, K, clusters, quad, assignment, stride, rem);
        }
        rem = (rem + 1) % stride;
    }
    free(active_grid);
    free(old_yx);
    for (int k = 0; k < K; k++) clusters[k].is_active = 1; /* preemptive.h:69-74 */
    assign_pass_bucketed(H, W, K, S, clusters, quad, lut, assignment, min_dists, 1, 0); /* :178-181 */
    if (quad_out) memcpy(quad_out, quad, (size_t)N * 4);
    if (precca_out) memcpy(precca_out, assignment, sizeof(uint16_t) * (size_t)N);
    memcpy(out, assignment, sizeof(uint16_t) * (size_t)N); /* :182-190 */
    int thres = (int)round((double)(S * S) * (double)min_size_factor); /* :16 */
    orc_enforce_connectivity(out, H, W, K, thres);                     /* :191-194 */
    free(quad); free(assignment); free(min_dists); free(lut);
}

/* the literal (unbucketed) scheduler, kept for tiny cases so the bucketed shortcut is itself checked */
void orc_assign_literal(int H, int W, int K, int S, OrcCluster* clusters, const uint8_t* quad, const uint16_t* lut,
                        uint16_t* assignment, uint16_t* min_dists, int stride, int rem) {
    assign_pass(H, W, K, S, clusters, quad, lut, assignment, min_dists, stride, rem);
}
void orc_assign_bucketed(int H, int W, int K, int S, OrcCluster* clusters, const uint8_t* quad, const uint16_t* lut,
                         uint16_t* assignment, uint16_t* min_dists, int stride, int rem) {
    assign_pass_bucketed(H, W, K, S, clusters, quad, lut, assignment, min_dists, stride, rem);
}
void orc_update(int H, int W, int K, OrcCluster* clusters, const uint8_t* quad, const uint16_t* assignment, int stride,
                int rem) {
    update_pass(H, W, K, clusters, quad, assignment, stride, rem);
}

/* ------------------------------------------------------------------ */
/* consumers of the label map: src/fast-slic.cpp:16-168                */
/* ------------------------------------------------------------------ */
#define ORC_MAX_CONN 12 /* fast-slic.cpp:17 */

/* fast_slic_get_connectivity, fast-slic.cpp:16-78.  The reference's K-word bit set only short-cuts the membership
 * test (a pair is linked <=> it is in both lists), so it is restated as the plain list search it guards.
 * counts[K], neighbors[K * 12]. */
void orc_get_connectivity(int H, int W, int K, const uint16_t* assignment, int32_t* counts, uint32_t* neighbors) {
    for (int k = 0; k < K; k++) counts[k] = 0;
    for (int i = 0; i < H - 1; i++) {
        for (int j = 0; j < W - 1; j++) {
            long base = (long)W * i + j;
            uint32_t source = assignment[base];
            if (source >= (uint32_t)K) continue;            /* :35 */
            int ns = counts[source];                        /* cached for the three probes, written back at :71 */
            const long probe[3] = {base + 1, base + W, base + W + 1}; /* :69-71 */
            for (int t = 0; t < 3; t++) {
                uint32_t target = assignment[probe[t]];
                if (target >= (uint32_t)K || source == target) continue;          /* :41 */
                int nt = counts[target];
                if (ns >= ORC_MAX_CONN || nt >= ORC_MAX_CONN) continue;           /* :43 */
                int exists = 0;
                for (int u = 0; u < ns && !exists; u++) exists = neighbors[source * ORC_MAX_CONN + u] == target; /* :47-52 */
                for (int u = 0; u < nt && !exists; u++) exists = neighbors[target * ORC_MAX_CONN + u] == source; /* :54-59 */
                if (exists) continue;
                neighbors[target * ORC_MAX_CONN + counts[target]++] = source;      /* :62 */
                neighbors[source * ORC_MAX_CONN + ns++] = target;                  /* :63 */
            }
            counts[source] = ns;                                                   /* :72 */
        }
    }
}

/* fast_slic_knn_connectivity (fast-slic.cpp:80-130) is NOT restated: it files every cluster under the cell index
 * (y / S) * nw + (x / S) evaluated in FLOAT (:88) -- 5.95 * 8 + 7.95 = 55 in a 6 x 8 grid -- so any centre in the lower
 * part of the last cell row indexes past the end of s_cells (heap overflow; the compiled reference segfaults on a plain
 * 120 x 160 / K = 48 seeding).  There is no defined behaviour to be identical to. */

/* fast_slic_get_mask_density / fast_slic_cluster_density_to_mask, fast-slic.cpp:141-168 */
void orc_get_mask_density(int H, int W, int K, const OrcCluster* clusters, const uint16_t* assignment, const uint8_t* mask,
                          uint8_t* densities) {
    int* sum = (int*)calloc((size_t)K, sizeof(int));
    for (long p = 0; p < (long)H * W; p++)
        if (assignment[p] < (uint16_t)K) sum[assignment[p]] += mask[p];
    for (int k = 0; k < K; k++) {
        unsigned den = clusters[k].num_members > 1u ? clusters[k].num_members : 1u;
        unsigned v = (unsigned)sum[k] / den;
        densities[k] = (uint8_t)(v < 255u ? v : 255u);
    }
    free(sum);
}
void orc_cluster_density_to_mask(int H, int W, int K, const uint16_t* assignment, const uint8_t* densities, uint8_t* result) {
    for (long p = 0; p < (long)H * W; p++) result[p] = assignment[p] < (uint16_t)K ? densities[assignment[p]] : 0;
}

/* ------------------------------------------------------------------ */
/* float-distance variants: src/context.cpp:394-499, context.h:100-125  */
/*   variant 0  ContextRealDist      BaseContext<float>: the default kernel with float min_dists and an         */
/*              UNtruncated float spatial patch  patch = coef * (|di| + |dj|)  (context.cpp:23-33, :259-298)      */
/*   variant 1  ContextRealDistL2    squared colour distance + patch = dj*dj + di*di, di = coef*(i-S) -- the      */
/*              reference's object code fuses it as fma(dj, dj, round(di*di)) (context.cpp:394-445)              */
/*   variant 2  ContextRealDistNoQ   float centroids, no window truncation to int16, per-pixel                   */
/*              |dr|+|dg|+|db|+|dx|+|dy| summed left to right, float division in the update (:447-499, :375-381)  */
/* Same scheduler, same strict '>' against min_dists (first visitor wins ties), same integer sums in update().    */
/* ------------------------------------------------------------------ */
static void assign_pass_real(int variant, int H, int W, int K, int S, OrcCluster* clusters, const uint8_t* quad, float coef,
                             uint16_t* assignment, float* min_dists, int stride, int rem) {
    for (long p = 0; p < (long)H * W; p++) min_dists[p] = 3.402823466e+38f; /* numeric_limits<float>::max(), :201-206 */
    for (int k = 0; k < K; k++) {
        float x = clusters[k].x, y = clusters[k].y;
        clusters[k].x = x < 0 ? 0 : (x > (float)(W - 1) ? (float)(W - 1) : x);
        clusters[k].y = y < 0 ? 0 : (y > (float)(H - 1) ? (float)(H - 1) : y);
    }
    int T = 2 * S + 32;
    int cell_W = ceil_int(W, T), cell_H = ceil_int(H, T);
    int ncell = cell_W * cell_H;
    int* start = (int*)calloc((size_t)ncell + 1, sizeof(int));
    int* items = (int*)malloc(sizeof(int) * (size_t)(K > 0 ? K : 1));
    for (int k = 0; k < K; k++) {
        if (!clusters[k].is_active) continue;
        start[cell_W * ((int)clusters[k].y / T) + ((int)clusters[k].x / T) + 1]++;
    }
    for (int c = 0; c < ncell; c++) start[c + 1] += start[c];
    int* fill = (int*)malloc(sizeof(int) * (size_t)(ncell > 0 ? ncell : 1));
    memcpy(fill, start, sizeof(int) * (size_t)ncell);
    for (int k = 0; k < K; k++) {
        if (!clusters[k].is_active) continue;
        items[fill[cell_W * ((int)clusters[k].y / T) + ((int)clusters[k].x / T)]++] = k;
    }
    for (int phase = 0; phase < 4; phase++)
        for (int ci = phase / 2; ci < cell_H; ci += 2)
            for (int cj = phase % 2; cj < cell_W; cj += 2) {
                int cell = ci * cell_W + cj;
                for (int t = start[cell]; t < start[cell + 1]; t++) {
                    int k = items[t];
                    const OrcCluster* c = &clusters[k];
                    int i0, i1, j0, j1;
                    int16_t cy = (int16_t)c->y, cx = (int16_t)c->x;
                    int16_t cr = (int16_t)c->r, cg = (int16_t)c->g, cb = (int16_t)c->b;
                    if (variant == 2) { /* :472-473: float arithmetic, truncated by my_max<int> / my_min<int> */
                        i0 = (int)(c->y - (float)S); if (i0 < 0) i0 = 0;
                        i1 = (int)(c->y + (float)S + 1.0f); if (i1 > H) i1 = H;
                        j0 = (int)(c->x - (float)S); if (j0 < 0) j0 = 0;
                        j1 = (int)(c->x + (float)S + 1.0f); if (j1 > W) j1 = W;
                        i1--; j1--;
                    } else {
                        i0 = cy - S < 0 ? 0 : cy - S; i1 = cy + S >= H ? H - 1 : cy + S;
                        j0 = cx - S < 0 ? 0 : cx - S; j1 = cx + S >= W ? W - 1 : cx + S;
                    }
                    for (int i = i0; i <= i1; i++) {
                        if (i % stride != rem) continue;
                        for (int j = j0; j <= j1; j++) {
                            long p = (long)i * W + j;
                            int r = quad[4 * p], g = quad[4 * p + 1], b = quad[4 * p + 2];
                            float d;
                            if (variant == 0) {
                                float patch = coef * (float)(abs(i - cy) + abs(j - cx));
                                d = patch + (float)(abs(r - cr) + abs(g - cg) + abs(b - cb));
                            } else if (variant == 1) {
                                float di = coef * (float)(i - cy), dj = coef * (float)(j - cx);
                                float patch = fmaf(dj, dj, di * di);
                                float dr = (float)(r - cr), dg = (float)(g - cg), db = (float)(b - cb);
                                d = patch + (dr * dr + dg * dg + db * db); /* integers below 2^24: exact in any order */
                            } else {
                                float dr = (float)r - c->r, dg = (float)g - c->g, db = (float)b - c->b;
                                float dy = coef * ((float)i - c->y), dx = coef * ((float)j - c->x);
                                d = fabsf(dr) + fabsf(dg) + fabsf(db) + fabsf(dx) + fabsf(dy);
                            }
                            if (min_dists[p] > d) {
                                min_dists[p] = d;
                                assignment[p] = c->number;
                            }
                        }
                    }
                }
            }
    free(start); free(items); free(fill);
}

static void update_pass_real(int variant, int H, int W, int K, OrcCluster* clusters, const uint8_t* quad,
                             const uint16_t* assignment, int stride, int rem) {
    if (variant != 2) {
        update_pass(H, W, K, clusters, quad, assignment, stride, rem);
        return;
    }
    int32_t* n = (int32_t*)calloc((size_t)K, sizeof(int32_t));
    int32_t* acc = (int32_t*)calloc((size_t)K * 5, sizeof(int32_t));
    for (int i = rem; i < H; i += stride)
        for (int j = 0; j < W; j++) {
            long p = (long)i * W + j;
            uint16_t c = assignment[p];
            if (c == 0xFFFF) continue;
            n[c]++;
            acc[5 * c + 0] += i; acc[5 * c + 1] += j;
            acc[5 * c + 2] += quad[4 * p]; acc[5 * c + 3] += quad[4 * p + 1]; acc[5 * c + 4] += quad[4 * p + 2];
        }
    for (int k = 0; k < K; k++) {
        OrcCluster* c = &clusters[k];
        if (!c->is_updatable) continue;
        c->num_members = (uint32_t)n[k];
        if (n[k] == 0) continue;
        c->y = (float)acc[5 * k + 0] / (float)n[k]; /* :375-381: (float)sum / int -> float division */
        c->x = (float)acc[5 * k + 1] / (float)n[k];
        c->r = (float)acc[5 * k + 2] / (float)n[k];
        c->g = (float)acc[5 * k + 3] / (float)n[k];
        c->b = (float)acc[5 * k + 4] / (float)n[k];
    }
    free(n); free(acc);
}

void orc_iterate_real(int variant, int H, int W, int K, const uint8_t* image, OrcCluster* clusters, uint16_t* out, int max_iter,
                      float compactness, float min_size_factor, int stride, int convert_to_lab, uint16_t* precca_out) {
    if (H <= 0 || W <= 0 || K <= 0) return;
    int S = (int16_t)sqrt((double)(H * W / K));
    long N = (long)H * W;
    uint8_t* quad = (uint8_t*)malloc((size_t)N * 4);
    uint16_t* assignment = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)N);
    float* min_dists = (float*)malloc(sizeof(float) * (size_t)N);
    int color_shift = convert_to_lab ? OUTPUT_SHIFT : 0;
    orc_rgb_to_quad(image, H, W, convert_to_lab, quad);
    for (int k = 0; k < K; k++) {
        int y = clampi((int)clusters[k].y, 0, H - 1), x = clampi((int)clusters[k].x, 0, W - 1);
        clusters[k].r = quad[4 * ((long)y * W + x)];
        clusters[k].g = quad[4 * ((long)y * W + x) + 1];
        clusters[k].b = quad[4 * ((long)y * W + x) + 2];
    }
    for (long p = 0; p < N; p++) assignment[p] = 0xFFFF;
    float coef = 1.0f / ((float)S / compactness);
    coef *= (float)(1 << color_shift);
    for (int k = 0; k < K; k++) clusters[k].is_updatable = 2;
    int rem = 0;
    for (int it = 0; it < max_iter; it++) {
        assign_pass_real(variant, H, W, K, S, clusters, quad, coef, assignment, min_dists, stride, rem);
        update_pass_real(variant, H, W, K, clusters, quad, assignment, stride, rem);
        rem = (rem + 1) % stride;
    }
    for (int k = 0; k < K; k++) clusters[k].is_active = 1;
    assign_pass_real(variant, H, W, K, S, clusters, quad, coef, assignment, min_dists, 1, 0);
    if (precca_out) memcpy(precca_out, assignment, sizeof(uint16_t) * (size_t)N);
    memcpy(out, assignment, sizeof(uint16_t) * (size_t)N);
    int thres = (int)round((double)(S * S) * (double)min_size_factor);
    orc_enforce_connectivity(out, H, W, K, thres);
    free(quad); free(assignment); free(min_dists);
}
