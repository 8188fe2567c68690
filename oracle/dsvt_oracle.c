/*
 * dsvt_oracle.c -- CPU restatement of the reference's hot-path algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker.  The product path (dsvt-ai-trt_b200/) never links or calls this file.
 *
 * Each function follows the cited reference kernel (jingyue202205/DSVT-AI-TRT @15b31c3) line by
 * line in plain scalar C with IEEE float32 arithmetic (build with -ffp-contract=off), executed in
 * SERIAL order: thread 0 first, then thread 1, ...  That serial order is the canonical outcome of
 * the reference's atomicAdd races and is what the CUDA path
 * reproduces deterministically.
 *
 * Parity pins: the reference has no tests; this oracle is pinned by (1) the known-answer values of
 * SURVEY.md Appendix B derived from the reference's own data/bin frames (5504 pillars / 454 sets /
 * 3558 max points -- constants that also appear in the reference's source comments), see
 * tests/test_oracle_kat.py, and (2) differential runs against the reference's own kernels compiled
 * unmodified into oracle/_ref (tests/test_reference_diff.py, GPU).  The set attention (TensorRT
 * internals) has no reference-side pin: PARITY UNPINNED for a3, cross-checked against
 * torch.nn.functional.multi_head_attention_forward instead.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- *
 * a1  points2Features.cu:669-705 (generateVoxels_random_kernel),
 *     :732-766 (generateBaseFeatures_kernel), :792-865 (generateFeatures_kernel)
 * ------------------------------------------------------------------------- */
ORACLE_API int oracle_points2features(
    const float* points, int n_points, int max_points, int max_rows, int max_pillars, int npv,
    float x_min, float x_max, float y_min, float y_max, float z_min, float z_max,
    float vx, float vy, float vz, int gx, int gy,
    float* point_features /*[max_rows,10]*/, int32_t* point_index_in_voxel /*[max_pillars,npv]*/,
    int32_t* coords /*[max_pillars,4]*/, int32_t* point_num_in_voxel /*[max_pillars]*/,
    int32_t* pillar_num, int32_t* point_num)
{
    const int G = gx * gy;
    int32_t* mask = (int32_t*) calloc((size_t) G, sizeof(int32_t));          /* :262 mask */
    int32_t* slots = (int32_t*) malloc((size_t) G * (size_t) npv * sizeof(int32_t)); /* stands for the dense voxels grid */
    if (!mask || !slots) { free(mask); free(slots); return -1; }
    memset(point_features, 0, (size_t) max_rows * 10 * sizeof(float));       /* :944-949 memsets */
    memset(point_index_in_voxel, 0, (size_t) max_pillars * npv * sizeof(int32_t));
    memset(coords, 0, (size_t) max_pillars * 4 * sizeof(int32_t));
    memset(point_num_in_voxel, 0, (size_t) max_pillars * sizeof(int32_t));

    if (n_points > max_points) n_points = max_points;                        /* :716-718 grid covers MAX_POINTS_NUM only */
    for (int i = 0; i < n_points; ++i) {                                     /* serial order == lowest indices win */
        const float x = points[i * 4 + 0], y = points[i * 4 + 1], z = points[i * 4 + 2];
        if (x < x_min || x >= x_max || y < y_min || y >= y_max || z < z_min || z >= z_max) continue; /* :683-685 */
        const int ix = (int) floorf((x - x_min) / vx);                       /* :687 */
        const int iy = (int) floorf((y - y_min) / vy);                       /* :688 */
        if (ix < 0 || ix >= gx || iy < 0 || iy >= gy) continue;              /* guard (the reference would write OOB) */
        const int cell = iy * gx + ix;                                       /* :689-690 */
        const int slot = mask[cell]++;                                       /* :697 atomicAdd */
        if (slot >= npv) continue;                                           /* :699 */
        slots[(size_t) cell * npv + slot] = i;                               /* :700-704 */
    }
    int V = 0, rows = 0;
    for (int cell = 0; cell < G; ++cell) {                                   /* canonical pillar order: ascending y*gx+x */
        int count = mask[cell];
        if (!(count > 0)) continue;                                          /* :747 */
        if (count > npv) count = npv;                                        /* :748 */
        const int pid = V++;                                                 /* :751 atomicAdd(pillar_num) */
        if (pid >= max_pillars) continue;                                    /* capacity guard (SURVEY A-5): dropped pillars emit no rows */
        if (count > max_rows - rows) count = max_rows - rows > 0 ? max_rows - rows : 0;
        point_num_in_voxel[pid] = count;                                     /* :753 */
        coords[pid * 4 + 0] = 0; coords[pid * 4 + 1] = 0;                    /* :755 */
        coords[pid * 4 + 2] = cell / gx; coords[pid * 4 + 3] = cell % gx;
        float cx = 0.f, cy = 0.f, cz = 0.f;                                  /* :805-807 */
        for (int s = 0; s < count; ++s) {                                    /* :809-818 sequential f32 sums */
            const float* p = points + (size_t) slots[(size_t) cell * npv + s] * 4;
            cx += p[0]; cy += p[1]; cz += p[2];
        }
        cx = cx / count; cy = cy / count; cz = cz / count;                   /* :819-821 */
        for (int s = 0; s < count; ++s) {
            const float* p = points + (size_t) slots[(size_t) cell * npv + s] * 4;
            const int row = rows + s;                                        /* :829 atomicAdd(point_num), canonical */
            point_index_in_voxel[(size_t) pid * npv + s] = row;              /* :830 */
            float* f = point_features + (size_t) row * 10;
            const float x = p[0], y = p[1], z = p[2];
            f[0] = x; f[1] = y; f[2] = z; f[3] = p[3];                       /* :838-841 */
            const int jx = (int) floorf((x - x_min) / vx);                   /* :844-846 */
            const int jy = (int) floorf((y - y_min) / vy);
            const int jz = (int) floorf((z - z_min) / vz);
            f[7] = (float) (x - ((jx + 0.5) * vx + x_min));                  /* :849-851 double arithmetic */
            f[8] = (float) (y - ((jy + 0.5) * vy + y_min));
            f[9] = (float) (z - ((jz + 0.5) * vz + z_min));
            f[4] = x - cx; f[5] = y - cy; f[6] = z - cz;                     /* :859-861 */
        }
        rows += (mask[cell] < npv ? mask[cell] : npv);
    }
    *pillar_num = V < max_pillars ? V : max_pillars;
    *point_num = rows < max_rows ? rows : max_rows;
    free(mask); free(slots);
    return 0;
}

/* ------------------------------------------------------------------------- *
 * windowPartition.cu:278-381 (splitWindow_kernel), host part :425-427
 * Canonical order: windows ascending dense index, voxels ascending id.
 * ------------------------------------------------------------------------- */
ORACLE_API int oracle_window_partition(
    const int32_t* coords, int voxel_num, int max_pillars, int max_win, int max_vpw,
    int sparse_x, int sparse_y, int sparse_z, int wx, int wy, int wz, int shx, int shy, int shz,
    int32_t* global_index /*[max_win,max_vpw]*/, int32_t* coors_in_win /*[max_win,max_vpw,3]*/,
    int32_t* voxel_num_in_win /*[max_win]*/, int32_t* win_num,
    int32_t* coors_2d /*[max_pillars,3]*/, float* coors_xy /*[max_pillars,2]*/)
{
    const int nwx = (int) (ceilf((float) (sparse_x / wx)) + 1);              /* :425 integer division inside */
    const int nwy = (int) (ceilf((float) (sparse_y / wy)) + 1);
    const int nwz = (int) (ceilf((float) (sparse_z / wz)) + 1);
    const int dn = nwx * nwy * nwz;
    int32_t* dense_count = (int32_t*) calloc((size_t) dn, sizeof(int32_t));
    int32_t* dense_slot = (int32_t*) malloc((size_t) dn * sizeof(int32_t));
    if (!dense_count || !dense_slot) { free(dense_count); free(dense_slot); return -1; }
    memset(global_index, 0, (size_t) max_win * max_vpw * sizeof(int32_t));
    memset(coors_in_win, 0, (size_t) max_win * max_vpw * 3 * sizeof(int32_t));
    memset(voxel_num_in_win, 0, (size_t) max_win * sizeof(int32_t));
    memset(coors_2d, 0, (size_t) max_pillars * 3 * sizeof(int32_t));
    memset(coors_xy, 0, (size_t) max_pillars * 2 * sizeof(float));
    if (voxel_num > max_pillars) voxel_num = max_pillars;
    for (int v = 0; v < voxel_num; ++v) {
        const uint32_t sx = (uint32_t) coords[v * 4 + 3] + shx, sy = (uint32_t) coords[v * 4 + 2] + shy,
                       sz = (uint32_t) coords[v * 4 + 1] + shz;              /* :290-292 */
        const uint32_t cx = sx / wx, cy = sy / wy, cz = sz / wz;             /* :294-296 */
        if (cx < (uint32_t) nwx && cy < (uint32_t) nwy && cz < (uint32_t) nwz)
            dense_count[cz * (nwy * nwx) + cy * nwx + cx]++;
    }
    int W = 0;
    for (int d = 0; d < dn; ++d) {
        dense_slot[d] = -1;
        if (dense_count[d] > 0) { if (W < max_win) dense_slot[d] = W; W++; }
        dense_count[d] = 0;
    }
    for (int v = 0; v < voxel_num; ++v) {
        const uint32_t sx = (uint32_t) coords[v * 4 + 3] + shx, sy = (uint32_t) coords[v * 4 + 2] + shy,
                       sz = (uint32_t) coords[v * 4 + 1] + shz;
        const uint32_t cx = sx / wx, cy = sy / wy, cz = sz / wz;
        const int ix = (int) (sx % wx), iy = (int) (sy % wy), iz = (int) (sz % wz);     /* :343-345 */
        coors_2d[v * 3 + 0] = iz; coors_2d[v * 3 + 1] = iy; coors_2d[v * 3 + 2] = ix;   /* :353-355 */
        coors_xy[v * 2 + 0] = (float) ix - (float) wx / 2;                              /* :358-359 */
        coors_xy[v * 2 + 1] = (float) iy - (float) wy / 2;
        if (!(cx < (uint32_t) nwx && cy < (uint32_t) nwy && cz < (uint32_t) nwz)) continue;
        const int d = (int) (cz * (nwy * nwx) + cy * nwx + cx);              /* :299 */
        const int slot = dense_slot[d];
        if (slot < 0) continue;
        const int pos = dense_count[d]++;                                    /* :302 atomicAdd, canonical = ascending v */
        if (pos >= max_vpw) continue;                                        /* :303 */
        voxel_num_in_win[slot] = pos + 1;                                    /* :334-339 */
        global_index[(size_t) slot * max_vpw + pos] = v;                     /* :342 */
        coors_in_win[((size_t) slot * max_vpw + pos) * 3 + 0] = iz;          /* :348-350 */
        coors_in_win[((size_t) slot * max_vpw + pos) * 3 + 1] = iy;
        coors_in_win[((size_t) slot * max_vpw + pos) * 3 + 2] = ix;
    }
    *win_num = W < max_win ? W : max_win;
    free(dense_count); free(dense_slot);
    return 0;
}

/* ------------------------------------------------------------------------- *
 * a2  getSet.cu:326-350 (getLocalIndex), :369-422 / :444-495 (sortY / sortX with the
 *     iterative quicksort :274-324), :517-567 (gather + mask), :589-609 (8-head expand)
 * ------------------------------------------------------------------------- */
static void oracle_swap(uint32_t* a, uint32_t* b) { uint32_t t = *a; *a = *b; *b = t; }

static int oracle_partition(uint32_t* arr, uint32_t* key, int l, int h) {   /* :274-292 (Lomuto) */
    const uint32_t x = key[h];
    int i = l - 1;
    for (int j = l; j <= h - 1; ++j) {
        if (key[j] <= x) { ++i; oracle_swap(&key[i], &key[j]); oracle_swap(&arr[i], &arr[j]); }
    }
    oracle_swap(&key[i + 1], &key[h]); oracle_swap(&arr[i + 1], &arr[h]);
    return i + 1;
}

static void oracle_quicksort(uint32_t* arr, uint32_t* key, int l, int h, int* stack) {  /* :294-324 */
    int top = -1;
    stack[++top] = l; stack[++top] = h;
    while (top >= 0) {
        h = stack[top--]; l = stack[top--];
        const int p = oracle_partition(arr, key, l, h);
        if (p - 1 > l) { stack[++top] = l; stack[++top] = p - 1; }
        if (p + 1 < h) { stack[++top] = p + 1; stack[++top] = h; }
    }
}

ORACLE_API int oracle_get_set(
    const int32_t* global_index, const int32_t* coors_in_win, const int32_t* voxel_num_in_win, int win_num,
    int S, int max_win, int max_vpw, int wx, int wy, int wz, int heads,
    int32_t* global_index_in_set /*[2,max_win,S]*/, float* set_voxel_mask /*[2,max_win,S]*/,
    int32_t* set_num, float* mask_expand_0 /*[max_win,heads,S]*/, float* mask_expand_1)
{
    uint32_t* sorted_y = (uint32_t*) malloc((size_t) max_vpw * sizeof(uint32_t));
    uint32_t* sorted_x = (uint32_t*) malloc((size_t) max_vpw * sizeof(uint32_t));
    uint32_t* key = (uint32_t*) malloc((size_t) max_vpw * sizeof(uint32_t));
    int* stack = (int*) malloc((size_t) (2 * max_vpw + 4) * sizeof(int));
    if (!sorted_y || !sorted_x || !key || !stack) { free(sorted_y); free(sorted_x); free(key); free(stack); return -1; }
    memset(global_index_in_set, 0, (size_t) 2 * max_win * S * sizeof(int32_t));       /* :680-686 memsets */
    memset(set_voxel_mask, 0, (size_t) 2 * max_win * S * sizeof(float));
    memset(mask_expand_0, 0, (size_t) max_win * heads * S * sizeof(float));
    memset(mask_expand_1, 0, (size_t) max_win * heads * S * sizeof(float));
    if (win_num > max_win) win_num = max_win;
    int nset = 0;
    for (int w = 0; w < win_num; ++w) {                                       /* canonical: ascending window slot */
        int N = voxel_num_in_win[w];
        if (N > max_vpw) N = max_vpw;
        if (N <= 0) continue;
        const int n_sets = (int) ceilf((float) N / S);                        /* :335 */
        const int base = nset;                                                /* :337 atomicAdd(set_num, n_sets) */
        nset += n_sets;
        const int32_t* gi = global_index + (size_t) w * max_vpw;
        const int32_t* cw = coors_in_win + (size_t) w * max_vpw * 3;
        for (int m = 0; m < N; ++m) {                                         /* :385-390 */
            sorted_y[m] = (uint32_t) gi[m];
            key[m] = (uint32_t) (cw[m * 3 + 1] * wx * wz + cw[m * 3 + 2] * wz + cw[m * 3 + 0]);
        }
        oracle_quicksort(sorted_y, key, 0, N - 1, stack);                     /* :421 */
        for (int m = 0; m < N; ++m) {                                         /* :460-465 */
            sorted_x[m] = (uint32_t) gi[m];
            key[m] = (uint32_t) (cw[m * 3 + 2] * wy * wz + cw[m * 3 + 1] * wz + cw[m * 3 + 0]);
        }
        oracle_quicksort(sorted_x, key, 0, N - 1, stack);                     /* :493 */
        for (int j = 0; j < n_sets; ++j) {
            const int set = base + j;
            if (set >= max_win) continue;                                     /* capacity guard */
            for (int k = 0; k < S; ++k) {
                const int local = (int) floorf((float) ((j * S + k) * N / S / n_sets));   /* :346 eq.(3) */
                const int32_t gy = (int32_t) sorted_y[local], gx = (int32_t) sorted_x[local];  /* :534-538 */
                global_index_in_set[((size_t) 0 * max_win + set) * S + k] = gy;
                global_index_in_set[((size_t) 1 * max_win + set) * S + k] = gx;
                float my = 0.f, mx = 0.f;
                if (k > 0) {                                                  /* :546-563 */
                    if (global_index_in_set[((size_t) 0 * max_win + set) * S + k - 1] == gy) my = (float) -3.4028235e+38;
                    if (global_index_in_set[((size_t) 1 * max_win + set) * S + k - 1] == gx) mx = (float) -3.4028235e+38;
                }
                set_voxel_mask[((size_t) 0 * max_win + set) * S + k] = my;
                set_voxel_mask[((size_t) 1 * max_win + set) * S + k] = mx;
                for (int h = 0; h < heads; ++h) {                             /* :598-606 */
                    mask_expand_0[((size_t) set * heads + h) * S + k] = my;
                    mask_expand_1[((size_t) set * heads + h) * S + k] = mx;
                }
            }
        }
    }
    *set_num = nset < max_win ? nset : max_win;
    free(sorted_y); free(sorted_x); free(key); free(stack);
    return 0;
}

/* ------------------------------------------------------------------------- *
 * a4  gelu.cu:201-211, constants params.h:75-77 (evaluated in double)
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_gelu(const float* x, int voxel_num, int max_pillars, int C, float* out)
{
    const double A = 0.5, B = 0.7978845608028654, Cc = 0.035677408136300125;
    memset(out, 0, (size_t) max_pillars * C * sizeof(float));                 /* :246 */
    if (voxel_num > max_pillars) voxel_num = max_pillars;
    const size_t n = (size_t) voxel_num * C;
    for (size_t i = 0; i < n; ++i) {
        const float v = x[i];
        out[i] = (float) ((A + A * tanh(v * (Cc * v * v + B))) * v);          /* :210 */
    }
}

/* ------------------------------------------------------------------------- *
 * a5  layerNorm.cu:297-309 (mean), :326-338 (biased variance), :261-279 (normalise)
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_layer_norm(const float* x, const float* residual, int voxel_num, int max_pillars, int C,
                                  const float* gamma, const float* beta, float eps, float* out)
{
    memset(out, 0, (size_t) max_pillars * C * sizeof(float));                 /* :395 */
    if (voxel_num > max_pillars) voxel_num = max_pillars;
    for (int r = 0; r < voxel_num; ++r) {
        const float* xr = x + (size_t) r * C;
        const float* rr = residual ? residual + (size_t) r * C : NULL;
        float avg = 0.0f;
        for (int j = 0; j < C; ++j) avg += rr ? xr[j] + rr[j] : xr[j];        /* :303-307 */
        const float mean = avg / C;
        float var = 0.0f;
        for (int j = 0; j < C; ++j) {                                         /* :332-336 */
            const float v = rr ? xr[j] + rr[j] : xr[j];
            var += (v - mean) * (v - mean);
        }
        var = var / C;
        for (int j = 0; j < C; ++j) {                                         /* :272-276 */
            const float v = rr ? xr[j] + rr[j] : xr[j];
            float t = (v - mean) / sqrtf(var + eps);
            t *= gamma[j];
            t += beta[j];
            out[(size_t) r * C + j] = t;
        }
    }
}

/* ------------------------------------------------------------------------- *
 * a6  filterBoxByScore.cu:266-309 -- kept candidates in ascending candidate index
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_filter_box(const float* scores, const int32_t* classes, const int32_t* xs, const int32_t* ys,
                                  const float* center, const float* center_z, const float* angle, const float* dim,
                                  int K, float x_min, float x_max, float y_min, float y_max, float z_min,
                                  float z_max, float vx, float vy, float score_thr,
                                  float* boxes /*[K,9]*/, int32_t* valid, int32_t* kept_index /*[K] or NULL*/)
{
    memset(boxes, 0, (size_t) K * 9 * sizeof(float));                         /* :364 */
    int n = 0;
    for (int i = 0; i < K; ++i) {                                             /* guarded: i < max_top_k (SURVEY A-10) */
        const float score = scores[i];
        float nx = (float) (uint32_t) xs[i] + center[i * 2 + 0];              /* :279-280 */
        float ny = (float) (uint32_t) ys[i] + center[i * 2 + 1];
        nx = fmaf(nx, vx, x_min);                                             /* :281-282, nvcc contracts to FMA */
        ny = fmaf(ny, vy, y_min);
        const float cz = center_z[i];
        if (!(nx >= x_min && nx < x_max && ny >= y_min && ny < y_max && cz >= z_min && cz < z_max)) continue; /* :287-291 */
        if (score >= score_thr) {                                             /* :293 */
            float* o = boxes + (size_t) n * 9;                                /* :295 atomicAdd(valid) */
            o[0] = nx; o[1] = ny; o[2] = cz;
            o[3] = dim[i * 3 + 0]; o[4] = dim[i * 3 + 1]; o[5] = dim[i * 3 + 2];
            o[6] = angle[i];
            o[7] = (float) (uint32_t) classes[i];                             /* :305 */
            o[8] = score;
            if (kept_index) kept_index[n] = i;
            ++n;
        }
    }
    *valid = n;
}

/* ------------------------------------------------------------------------- *
 * getValueByIndex.cu:282-303 and mapSetFeature2voxel.cu:258-275
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_get_value_by_index(const float* x, const float* pos, const int32_t* idx /*[2,max_sets,S]*/,
                                          int set_num, int max_sets, int S, int C, int axis,
                                          float* q, float* k, float* v)
{
    const size_t frame = (size_t) max_sets * S * C;
    memset(q, 0, frame * sizeof(float)); memset(k, 0, frame * sizeof(float)); memset(v, 0, frame * sizeof(float));
    if (set_num > max_sets) set_num = max_sets;
    const int32_t* id = idx + (size_t) axis * max_sets * S;
    for (int t = 0; t < set_num * S; ++t) {
        const int g = id[t];
        for (int c = 0; c < C; ++c) {
            const float a = x[(size_t) g * C + c], p = pos[(size_t) g * C + c];
            q[(size_t) t * C + c] = a + p; k[(size_t) t * C + c] = a + p; v[(size_t) t * C + c] = a;
        }
    }
}

ORACLE_API void oracle_map_set_feature2voxel(const float* feat, const int32_t* idx, int set_num, int max_sets, int S,
                                             int C, int axis, int max_pillars, float* out)
{
    memset(out, 0, (size_t) max_pillars * C * sizeof(float));
    if (set_num > max_sets) set_num = max_sets;
    const int32_t* id = idx + (size_t) axis * max_sets * S;
    for (int t = 0; t < set_num * S; ++t)
        memcpy(out + (size_t) id[t] * C, feat + (size_t) t * C, (size_t) C * sizeof(float));
}

/* ------------------------------------------------------------------------- *
 * a3  multHeadAttention(), src/dsvt-ai-trt.cpp:288-458.  Arithmetic executes inside TensorRT
 *     in the reference; this follows the layer sequence: three FCs (:328-330), q / sqrt(C/heads)
 *     (:386-405), bmm (:410), + key mask broadcast over query rows (:412), softmax over keys
 *     (:414-415), bmm (:417), head-major concat, out FC (:448).  Weights in PyTorch [out][in].
 * ------------------------------------------------------------------------- */
ORACLE_API int oracle_set_attention(const float* q, const float* k, const float* v, const float* mask /*[sets,heads,S]*/,
                                    int n_sets, int S, int C, int heads,
                                    const float* w_in /*[3C,C]*/, const float* b_in /*[3C]*/,
                                    const float* w_out /*[C,C]*/, const float* b_out /*[C]*/,
                                    float* out /*[sets,S,C]*/)
{
    const int D = C / heads;                                                  /* integer division (:359) */
    const float scale = (float) sqrt((double) (C / heads));                   /* :386 */
    float* Q = (float*) malloc((size_t) S * C * sizeof(float));
    float* K = (float*) malloc((size_t) S * C * sizeof(float));
    float* V = (float*) malloc((size_t) S * C * sizeof(float));
    float* O = (float*) malloc((size_t) S * C * sizeof(float));
    float* P = (float*) malloc((size_t) S * sizeof(float));
    if (!Q || !K || !V || !O || !P) { free(Q); free(K); free(V); free(O); free(P); return -1; }
    for (int s = 0; s < n_sets; ++s) {
        const float* qs = q + (size_t) s * S * C; const float* ks = k + (size_t) s * S * C;
        const float* vs = v + (size_t) s * S * C;
        for (int t = 0; t < S; ++t) {
            for (int n = 0; n < C; ++n) {
                float aq = 0.f, ak = 0.f, av = 0.f;
                const float* wq = w_in + (size_t) n * C; const float* wk = w_in + (size_t) (C + n) * C;
                const float* wv = w_in + (size_t) (2 * C + n) * C;
                for (int c = 0; c < C; ++c) {
                    aq += qs[t * C + c] * wq[c]; ak += ks[t * C + c] * wk[c]; av += vs[t * C + c] * wv[c];
                }
                Q[t * C + n] = (aq + b_in[n]) / scale;                        /* :405 kDIV */
                K[t * C + n] = ak + b_in[C + n];
                V[t * C + n] = av + b_in[2 * C + n];
            }
        }
        for (int h = 0; h < heads; ++h) {
            const float* mk = mask + ((size_t) s * heads + h) * S;
            for (int i = 0; i < S; ++i) {
                float mx = -INFINITY;
                for (int j = 0; j < S; ++j) {
                    float a = 0.f;
                    for (int d = 0; d < D; ++d) a += Q[i * C + h * D + d] * K[j * C + h * D + d];
                    a += mk[j];                                               /* :412 */
                    P[j] = a;
                    if (a > mx) mx = a;
                }
                float sum = 0.f;
                for (int j = 0; j < S; ++j) { P[j] = expf(P[j] - mx); sum += P[j]; }
                for (int j = 0; j < S; ++j) P[j] /= sum;
                for (int d = 0; d < D; ++d) {
                    float a = 0.f;
                    for (int j = 0; j < S; ++j) a += P[j] * V[j * C + h * D + d];
                    O[i * C + h * D + d] = a;                                 /* head-major concat (:419-446) */
                }
            }
        }
        for (int t = 0; t < S; ++t) {
            for (int n = 0; n < C; ++n) {
                float a = 0.f;
                const float* wo = w_out + (size_t) n * C;
                for (int c = 0; c < C; ++c) a += O[t * C + c] * wo[c];
                out[((size_t) s * S + t) * C + n] = a + b_out[n];             /* :448 */
            }
        }
    }
    free(Q); free(K); free(V); free(O); free(P);
    return 0;
}

ORACLE_API int oracle_abi_version(void) { return 1; }

/* ------------------------------------------------------------------------- *
 * next #3  torchScatterMax.cu:201-262 (generateMax_kernel) after the memsets of :300-301
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_torch_scatter_max(const float* point_features, const int* point_index_in_voxel,
                                         const int* point_num_in_voxel, int voxel_num, int max_points, int max_pillars,
                                         int npv, int F, float* max_point_features, float* max_voxel_features)
{
    memset(max_point_features, 0, (size_t) max_points * F * sizeof(float));        /* :300 */
    memset(max_voxel_features, 0, (size_t) max_pillars * F * sizeof(float));       /* :301 */
    if (voxel_num > max_pillars) voxel_num = max_pillars;
    for (int v = 0; v < voxel_num; ++v) {
        const int* idx = point_index_in_voxel + (size_t) v * npv;
        const int n = point_num_in_voxel[v];
        for (int k = 0; k < F; ++k) {
            float m = -1000000.0f;                                                 /* :216 */
            for (int i = 0; i < n; ++i) {
                const float f = point_features[(size_t) idx[i] * F + k];
                if (f > m) m = f;                                                  /* :230-238 */
            }
            max_voxel_features[(size_t) v * F + k] = m;                            /* :245 */
            for (int i = 0; i < n; ++i) max_point_features[(size_t) idx[i] * F + k] = m;   /* :251-259 */
        }
    }
}

/* ------------------------------------------------------------------------- *
 * next #3  map2bev.cu:250-265 (Map2Bev_kernel) after the memset of :305
 * ------------------------------------------------------------------------- */
ORACLE_API void oracle_map2bev(const float* voxel_features, const int* coords, int voxel_num, int max_pillars, int C,
                               int gx, int gy, float* map)
{
    memset(map, 0, (size_t) gx * gy * C * sizeof(float));                          /* :305 */
    if (voxel_num > max_pillars) voxel_num = max_pillars;
    for (int v = 0; v < voxel_num; ++v) {
        const unsigned y = (unsigned) coords[(size_t) v * 4 + 2], x = (unsigned) coords[(size_t) v * 4 + 3];   /* :258-260 */
        if (y >= (unsigned) gy || x >= (unsigned) gx) continue;                    /* the reference writes out of bounds */
        for (int c = 0; c < C; ++c)
            map[((size_t) y * gx + x) * C + c] = voxel_features[(size_t) v * C + c];   /* :263 */
    }
}

/* ------------------------------------------------------------------------- *
 * next #4 tail  rotated NMS on the host: include/helper.h:92-283 (nms_cpu, box_overlap, intersection, check_box2d,
 * rotate_around_center; the code the reference took from CUDA-PointPillars' postprocess.cpp).  float arithmetic and
 * the order of operations are kept; std::sort is restated as a STABLE descending sort (ties keep input order).
 * boxes [n,9] = (x, y, z, dx, dy, dz, angle, class, score) as FilterBoxByScorePlugin emits them (:1954).
 * keep[] receives the input indices of the surviving boxes in output order; returns their number.
 * ------------------------------------------------------------------------- */
typedef struct { float x, y; } nms_f2;
typedef struct { float x, y, z, w, l, h, rt; int id; float score; } nms_box;
static const float kNmsEps = 1e-8f;                                                /* helper.h:26 ThresHold */

static float nms_cross(nms_f2 p1, nms_f2 p2, nms_f2 p0) {                         /* :107-109 */
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static int nms_check_box2d(const nms_box* box, nms_f2 p) {                         /* :111-121 */
    const float MARGIN = 1e-2f;
    const float angle_cos = cosf(-box->rt), angle_sin = sinf(-box->rt);            /* cos/sin(float) -> float overloads */
    const float rot_x = (p.x - box->x) * angle_cos + (p.y - box->y) * (-angle_sin);
    const float rot_y = (p.x - box->x) * angle_sin + (p.y - box->y) * angle_cos;
    return (fabsf(rot_x) < box->w / 2 + MARGIN && fabsf(rot_y) < box->l / 2 + MARGIN);
}
static int nms_intersection(nms_f2 p1, nms_f2 p0, nms_f2 q1, nms_f2 q0, nms_f2* ans) {   /* :123-157 */
    if ((fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
         fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)) == 0)
        return 0;
    const float s1 = nms_cross(q0, p1, p0), s2 = nms_cross(p1, q1, p0);
    const float s3 = nms_cross(p0, q1, q0), s4 = nms_cross(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    const float s5 = nms_cross(q1, p1, p0);
    if (fabsf(s5 - s1) > kNmsEps) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}
static void nms_rotate(nms_f2 c, float ac, float as, nms_f2* p) {                  /* :159-164 */
    const float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
    const float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
    p->x = nx; p->y = ny;
}
static float nms_box_overlap(const nms_box* a, const nms_box* b) {                 /* :166-255 */
    const float a_dx = a->w / 2, b_dx = b->w / 2, a_dy = a->l / 2, b_dy = b->l / 2;
    nms_f2 ca[5] = {{a->x - a_dx, a->y - a_dy}, {a->x + a_dx, a->y - a_dy}, {a->x + a_dx, a->y + a_dy}, {a->x - a_dx, a->y + a_dy}, {0, 0}};
    nms_f2 cb[5] = {{b->x - b_dx, b->y - b_dy}, {b->x + b_dx, b->y - b_dy}, {b->x + b_dx, b->y + b_dy}, {b->x - b_dx, b->y + b_dy}, {0, 0}};
    const nms_f2 center_a = {a->x, a->y}, center_b = {b->x, b->y};
    nms_f2 pts[16], pc = {0, 0};
    int cnt = 0;
    const float a_cos = cosf(a->rt), a_sin = sinf(a->rt), b_cos = cosf(b->rt), b_sin = sinf(b->rt);
    for (int k = 0; k < 4; ++k) { nms_rotate(center_a, a_cos, a_sin, &ca[k]); nms_rotate(center_b, b_cos, b_sin, &cb[k]); }
    ca[4] = ca[0]; cb[4] = cb[0];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (nms_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) { pc.x += pts[cnt].x; pc.y += pts[cnt].y; ++cnt; }
    for (int k = 0; k < 4; ++k) {
        if (nms_check_box2d(a, cb[k])) { pc.x += cb[k].x; pc.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (nms_check_box2d(b, ca[k])) { pc.x += ca[k].x; pc.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    pc.x /= cnt; pc.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - pc.y, pts[i].x - pc.x) > atan2f(pts[i + 1].y - pc.y, pts[i + 1].x - pc.x)) {
                const nms_f2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0;
    for (int k = 0; k < cnt - 1; ++k) {
        const nms_f2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
        area += (u.x * v.y - u.y * v.x);
    }
    return (float) (fabs((double) area) / 2.0);                                    /* fabs(area) / 2.0 in double, returned as float */
}
/* pairwise IoU of boxes i, j (for the tests: distance of every decision from the threshold) */
ORACLE_API float oracle_nms_iou(const float* boxes, int i, int j)
{
    nms_box a = {boxes[i * 9], boxes[i * 9 + 1], boxes[i * 9 + 2], boxes[i * 9 + 3], boxes[i * 9 + 4], boxes[i * 9 + 5], boxes[i * 9 + 6], i, boxes[i * 9 + 8]};
    nms_box b = {boxes[j * 9], boxes[j * 9 + 1], boxes[j * 9 + 2], boxes[j * 9 + 3], boxes[j * 9 + 4], boxes[j * 9 + 5], boxes[j * 9 + 6], j, boxes[j * 9 + 8]};
    const float sa = a.w * a.l, sb = b.w * b.l, so = nms_box_overlap(&a, &b);
    return so / fmaxf(sa + sb - so, kNmsEps);
}
ORACLE_API int oracle_nms(const float* boxes, int n, float nms_thresh, int* keep)  /* :257-283 */
{
    if (n <= 0) return 0;
    int* order = (int*) malloc((size_t) n * sizeof(int));
    char* suppressed = (char*) calloc((size_t) n, 1);
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i) {                       /* stable insertion sort, descending score */
        const int v = order[i];
        int j = i - 1;
        while (j >= 0 && boxes[order[j] * 9 + 8] < boxes[v * 9 + 8]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = v;
    }
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        if (suppressed[i]) continue;
        keep[kept++] = order[i];
        for (int j = i + 1; j < n; ++j) {
            if (suppressed[j]) continue;
            if (oracle_nms_iou(boxes, order[i], order[j]) >= nms_thresh) suppressed[j] = 1;
        }
    }
    free(order); free(suppressed);
    return kept;
}
