// Fused VFE (included by attention_split.cu inside its anonymous namespace): the two PFN layers and the two per-pillar
// maxima of the reference's pillar feature net (src/dsvt-ai-trt.cpp:571-590) in ONE kernel,
//   h0 = relu(bn0(x W0^T))            [Pc, 96]    Linear(10 -> 96) + folded BatchNorm + ReLU          (:268-286, :577)
//   m0 = max over the pillar's rows   [V, 96]     TorchScatterMaxPlugin 0                             (:579)
//   h1 = relu([h0 | m0(pillar)] W1^T) [Pc, 192]   concat (:583-587) + Linear(192 -> 192) + BN + ReLU
//   out = max over the pillar's rows  [V, 192]    TorchScatterMaxPlugin 1, output 1 -- the only tensor the graph reads (:589)
// so that none of the four per-point tensors (h0, the per-point copies of m0, h1, the per-point copies of out: ~0.9 GB per
// frame at Waymo size) ever reaches memory: the kernel reads the voxeliser's point rows (40 B each) and writes one 768-byte
// row per pillar.  The voxeliser emits rows pillar-major, a pillar's rows consecutive (point_index_in_voxel[v][i] = first
// row + i, at most max_num_points_per_voxel <= 64 of them), so a tile is "the pillars whose first row lies in
// [80 t, 80 t + 80)": at most 80 + 63 rows... the tile kernel's M is 128, so the window is 128 - 63 = 64 rows when npv = 64 and
// 80 rows for npv <= 48 (kVfeWin below is chosen by the launch).  Every pillar belongs to exactly one tile: no atomics.
//
// Arithmetic = the separate kernels': small_linear_kernel<10> for h0 (sequential FMA chain, Scale, ReLU), strict '>' maxima
// from -1000000 (torchScatterMax.cu:216,:230), the tile GEMM's products / order / epilogue for h1.
constexpr int kVfeH0Stride = 97;                 // floats per staged row (odd: conflict-free row- and column-wise)
constexpr int kVfeMaxPillars = 80;               // pillars per tile (every pillar has >= 1 row)
constexpr int kVfeA = kNumK * 2 * kATerm;        // 98304 B: [h0 | m0] as the GEMM's A image
constexpr int kVfeM0 = kVfeMaxPillars * kVfeH0Stride * 4;       // 31040 B
constexpr int kVfeH0 = kBM * kVfeH0Stride * 4;                  // 49664 B  (>= the 2 x 24 KB weight ring that reuses it)
constexpr int kVfeSmem = kVfeA + kVfeM0 + kVfeH0;               // 179008 B
static_assert(kVfeH0 >= 2 * 2 * kBTerm, "weight ring aliases the h0 staging area");
static_assert(kVfeA + kVfeM0 >= kBM * kLnStride * 4, "finished tile aliases the A image");
constexpr int kVfeWorkerWarps = 16, kVfeWorkers = kVfeWorkerWarps * 32, kVfeThreads = kVfeWorkers + 64;

struct VfeArgs {
    const float* x;           // [max_points, 10] point rows (Points2FeaturesPlugin output 0)
    const int* piv;           // [max_pillars, npv] row ids per pillar (output 1)
    const float* pfn0;        // dsvt_small_linear blob: W0 [96][10] | scale [96] | shift [96]
    const uint8_t* w1_img;    // PFN layer 1 as the tile GEMM's block image (6 chunks of 24 KB), BatchNorm folded
    const float* bias1;       // [192]
    float out_mul1;
    float* out;               // [max_pillars, 192]
    const int* tiles;         // [0] = tile count, [1 + t] = first pillar of tile t, [1 + count] = V
    int npv, win;
};

// tile plan: one thread per pillar; also zero-fills the output rows beyond the pillar count (the plugin's contract)
__global__ void __launch_bounds__(256)
vfe_plan_kernel(const int* __restrict__ piv, const int* __restrict__ voxel_num, int* __restrict__ tiles, float4* __restrict__ out,
                int max_pillars, int npv, int win, int zero_tails)
{
    int V = voxel_num[0];
    V = V < max_pillars ? V : max_pillars;
    const int stride = gridDim.x * blockDim.x;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < max_pillars; v += stride) {
        if (v < V) {
            const int t = __ldg(piv + (size_t) v * npv) / win;
            const int tp = v > 0 ? __ldg(piv + (size_t) (v - 1) * npv) / win : -1;
            for (int tt = tp + 1; tt <= t; ++tt) tiles[1 + tt] = v;       // windows without a pillar start (npv > win): empty tiles
            if (v == V - 1) { tiles[0] = t + 1; tiles[2 + t] = V; }
        } else if (zero_tails) {
            for (int c = 0; c < kC / 4; ++c) stg_zero4(out + (size_t) v * (kC / 4) + c);
        }
    }
    if (V == 0 && blockIdx.x == 0 && threadIdx.x == 0) tiles[0] = 0;
}

__global__ void __launch_bounds__(kVfeThreads, 1)
vfe_fused_kernel(const __grid_constant__ VfeArgs a, const int* __restrict__ voxel_num, const int* __restrict__ point_num,
                 int max_points, int max_pillars)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full, h0_dead, w_full[2], w_empty[2], acc_full;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_w0[96 * 12];      // per column n: W0[n][0..9] | scale | shift (three 16-byte reads)
    __shared__ int s_first[kVfeMaxPillars + 1];
    __shared__ unsigned char s_pil[kBM];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x;
    if (t >= __ldg(a.tiles)) return;
    int V = voxel_num[0], Pc = point_num[0];
    V = V < max_pillars ? V : max_pillars;
    Pc = Pc < max_points ? Pc : max_points;
    const int pa = __ldg(a.tiles + 1 + t), pb = __ldg(a.tiles + 2 + t);
    int npil = pb - pa;
    npil = npil < kVfeMaxPillars ? npil : kVfeMaxPillars;          // (guard: cannot exceed the window by construction)
    if (npil <= 0) return;
    const int r0 = __ldg(a.piv + (size_t) pa * a.npv);

    float* A_img = reinterpret_cast<float*>(smem);                 // as bytes below
    float* M0 = reinterpret_cast<float*>(smem + kVfeA);
    float* H0 = reinterpret_cast<float*>(smem + kVfeA + kVfeM0);
    uint8_t* wring = smem + kVfeA + kVfeM0;
    (void) A_img;

    if (tid == 0) TP(0);
    if (tid == 0) {
        mbar_init(&a_full, kVfeWorkers); mbar_init(&h0_dead, kVfeWorkers); mbar_init(&acc_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == kVfeWorkerWarps) tmem_alloc<256>(&tmem_slot);
    for (int i = tid; i < 96 * 12; i += kVfeThreads) {
        const int n = i / 12, k = i - n * 12;
        s_w0[i] = __ldg(a.pfn0 + (k < 10 ? n * 10 + k : (k == 10 ? 960 + n : 1056 + n)));
    }
    for (int j = tid; j <= npil; j += kVfeThreads) {
        const int p = pa + j;
        int fr = p < V ? __ldg(a.piv + (size_t) p * a.npv) : Pc;
        if (j == npil && pb >= V) fr = Pc;
        s_first[j] = fr - r0;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    int nrows = s_first[npil];
    nrows = nrows < kBM ? nrows : kBM;                             // (guard)
    if (tid == 0) TP(1);

    if (warp < kVfeWorkerWarps) {
        for (int j = tid; j < npil; j += kVfeWorkers)
            for (int r = s_first[j]; r < s_first[j + 1] && r < kBM; ++r) s_pil[r] = (unsigned char) j;
        // ---- step A: h0 = relu(bn0(x W0^T)) -> FP32 staging (for the maxima) + columns 0..95 of the A image ------------------
        {
            const int row = tid & 127, cg = tid >> 7;              // 24 columns = 3 K pieces of 8 per thread
            float xin[10];
            if (row < nrows) {
                const float2* xp = reinterpret_cast<const float2*>(a.x + (size_t) (r0 + row) * 10);
#pragma unroll
                for (int k = 0; k < 5; ++k) { const float2 v = __ldg(xp + k); xin[2 * k] = v.x; xin[2 * k + 1] = v.y; }
            } else {
#pragma unroll
                for (int k = 0; k < 10; ++k) xin[k] = 0.f;
            }
#pragma unroll
            for (int p8 = 0; p8 < 3; ++p8) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int n = cg * 24 + p8 * 8 + e;
                    const float4 w0 = *reinterpret_cast<const float4*>(s_w0 + n * 12), w1 = *reinterpret_cast<const float4*>(s_w0 + n * 12 + 4),
                                 w2 = *reinterpret_cast<const float4*>(s_w0 + n * 12 + 8);
                    float acc = 0.f;
                    acc = fmaf(xin[0], w0.x, acc); acc = fmaf(xin[1], w0.y, acc); acc = fmaf(xin[2], w0.z, acc); acc = fmaf(xin[3], w0.w, acc);
                    acc = fmaf(xin[4], w1.x, acc); acc = fmaf(xin[5], w1.y, acc); acc = fmaf(xin[6], w1.z, acc); acc = fmaf(xin[7], w1.w, acc);
                    acc = fmaf(xin[8], w2.x, acc); acc = fmaf(xin[9], w2.y, acc);
                    acc = fmaf(acc, w2.z, w2.w);
                    v[e] = row < nrows ? fmaxf(acc, 0.f) : 0.f;
                    H0[row * kVfeH0Stride + n] = v[e];
                    v[e] = v[e] + 0.f;                             // the tile GEMM's producers add the (absent) second operand
                }
                const uint4 hi = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                const float2 h0 = unpack_h2(hi.x), h1 = unpack_h2(hi.y), h2 = unpack_h2(hi.z), h3 = unpack_h2(hi.w);
                const uint4 lo = make_uint4(pack_h2(v[0] - h0.x, v[1] - h0.y), pack_h2(v[2] - h1.x, v[3] - h1.y),
                                            pack_h2(v[4] - h2.x, v[5] - h2.y), pack_h2(v[6] - h3.x, v[7] - h3.y));
                const int piece = cg * 3 + p8, kc = piece >> 2, c16 = piece & 3;
                uint8_t* chunk = smem + kc * (2 * kATerm);
                *reinterpret_cast<uint4*>(chunk + c16 * (kBM * 16) + row * 16) = hi;
                *reinterpret_cast<uint4*>(chunk + kATerm + c16 * (kBM * 16) + row * 16) = lo;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kVfeWorkers) : "memory");
        if (tid == 0) TP(2);
        // ---- step B: m0 = per-pillar maxima of h0 -------------------------------------------------------------------------------
        if (tid < 480) {
            const int c = tid % 96;
            for (int j = tid / 96; j < npil; j += 10) {        // two pillars per iteration: independent load / compare chains
                const int j2 = j + 5 < npil ? j + 5 : j;
                float m = -1000000.0f, m2 = -1000000.0f;
                const int rb = s_first[j], re = s_first[j + 1] < kBM ? s_first[j + 1] : kBM;
                const int rb2 = s_first[j2], re2 = s_first[j2 + 1] < kBM ? s_first[j2 + 1] : kBM;
                const int n = re - rb > re2 - rb2 ? re - rb : re2 - rb2;
                for (int i = 0; i < n; ++i) {
                    const int r = rb + i < re ? rb + i : rb, r2 = rb2 + i < re2 ? rb2 + i : rb2;      // clamped: a repeated row
                    const float v = H0[r * kVfeH0Stride + c], v2 = H0[r2 * kVfeH0Stride + c];          // does not change the max
                    m = v > m ? v : m; m2 = v2 > m2 ? v2 : m2;
                }
                M0[j * kVfeH0Stride + c] = m;
                M0[j2 * kVfeH0Stride + c] = m2;
            }
        }
        mbar_arrive(&h0_dead);                                     // the weight ring may overwrite the h0 staging area
        asm volatile("bar.sync 1, %0;" ::"n"(kVfeWorkers) : "memory");
        if (tid == 0) TP(3);
        // ---- step C: columns 96..191 of the A image = the pillar's maxima, per row ----------------------------------------------
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int task = i * kVfeWorkers + tid, row = task & 127, piece = 12 + (task >> 7);
            float v[8];
            if (row < nrows) {
                const float* mp = M0 + (int) s_pil[row] * kVfeH0Stride + (piece - 12) * 8;
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = mp[e] + 0.f;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = 0.f;
            }
            const uint4 hi = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            const float2 h0 = unpack_h2(hi.x), h1 = unpack_h2(hi.y), h2 = unpack_h2(hi.z), h3 = unpack_h2(hi.w);
            const uint4 lo = make_uint4(pack_h2(v[0] - h0.x, v[1] - h0.y), pack_h2(v[2] - h1.x, v[3] - h1.y),
                                        pack_h2(v[4] - h2.x, v[5] - h2.y), pack_h2(v[6] - h3.x, v[7] - h3.y));
            const int kc = piece >> 2, c16 = piece & 3;
            uint8_t* chunk = smem + kc * (2 * kATerm);
            *reinterpret_cast<uint4*>(chunk + c16 * (kBM * 16) + row * 16) = hi;
            *reinterpret_cast<uint4*>(chunk + kATerm + c16 * (kBM * 16) + row * 16) = lo;
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_full);
        if (tid == 0) TP(4);
        // ---- step E: h1 = relu(acc * out_mul + bias) -> finished tile -> per-pillar maxima -> one row per pillar ---------------
        mbar_wait(&acc_full, 0);
        tc_fence_after_sync();
        if (tid == 0) TP(9);
        float* tile = reinterpret_cast<float*>(smem);
        {
            constexpr int kCols = kBN / (kVfeWorkerWarps / 4);     // 48 accumulator columns per warp
            const int q4 = warp & 3, hf = warp >> 2;
            const uint32_t tlane = tmem + ((uint32_t) (q4 * 32) << 16) + hf * kCols;
            float* trow = tile + (size_t) (q4 * 32 + lane) * kLnStride + hf * kCols;
#pragma unroll 1
            for (int j0 = 0; j0 < kCols; j0 += 16) {
                uint32_t r[16];
                tmem_ld16(tlane + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias1 + hf * kCols + j0 + 4 * j));
                    float4 y = make_float4((__uint_as_float(r[4 * j]) * a.out_mul1 + bb.x) * 1.0f, (__uint_as_float(r[4 * j + 1]) * a.out_mul1 + bb.y) * 1.0f,
                                           (__uint_as_float(r[4 * j + 2]) * a.out_mul1 + bb.z) * 1.0f, (__uint_as_float(r[4 * j + 3]) * a.out_mul1 + bb.w) * 1.0f);
                    y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
                    *reinterpret_cast<float4*>(trow + j0 + 4 * j) = y;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kVfeWorkers) : "memory");
        if (tid == 0) TP(10);
        for (int task = tid; task < npil * (kC / 4); task += kVfeWorkers) {
            const int j = task / (kC / 4), c4 = task - j * (kC / 4);
            float4 m = make_float4(-1000000.0f, -1000000.0f, -1000000.0f, -1000000.0f);
            const int re = s_first[j + 1] < kBM ? s_first[j + 1] : kBM;
            for (int r = s_first[j]; r < re; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(tile + (size_t) r * kLnStride + c4 * 4);
                m.x = v.x > m.x ? v.x : m.x; m.y = v.y > m.y ? v.y : m.y; m.z = v.z > m.z ? v.z : m.z; m.w = v.w > m.w ? v.w : m.w;
            }
            reinterpret_cast<float4*>(a.out + (size_t) (pa + j) * kC)[c4] = m;
        }
    } else if (warp == kVfeWorkerWarps) {
        // =========================== MMA ISSUE (converged warp) ==========================================
        const uint32_t idesc = make_idesc(kFmtF16, kBM, kBN);
        const uint32_t sbase = smem_u32(smem), wbase = smem_u32(wring);
        const uint64_t a_base = make_smem_desc(sbase, kBM * 16, 128), w_base = make_smem_desc(wbase, kBN * 16, 128);
        mbar_wait(&a_full, 0); __syncwarp();
        tc_fence_after_sync();
#pragma unroll 1
        for (int kc = 0; kc < kNumK; ++kc) {
            const int ws = kc & 1;
            mbar_wait(&w_full[ws], (kc >> 1) & 1); __syncwarp();
            tc_fence_after_sync();
            if (lane == 0) TP(14 + kc);
            const uint64_t ad = a_base + (uint64_t) ((kc * 2 * kATerm) >> 4), wd = w_base + (uint64_t) ((ws * 2 * kBTerm) >> 4);
#pragma unroll
            for (int ks = 0; ks < kBK / 16; ++ks) {
                const uint64_t a_hi = ad + (uint64_t) ((ks * 2 * (kBM * 16)) >> 4), a_lo = a_hi + (uint64_t) (kATerm >> 4);
                const uint64_t b_hi = wd + (uint64_t) ((ks * 2 * (kBN * 16)) >> 4), b_lo = b_hi + (uint64_t) (kBTerm >> 4);
                umma_f16_w(tmem, a_lo, b_hi, idesc, (kc | ks) != 0);
                umma_f16_w(tmem, a_hi, b_lo, idesc, 1);
                umma_f16_w(tmem, a_hi, b_hi, idesc, 1);
            }
            umma_commit_w(&w_empty[ws]);
        }
        umma_commit_w(&acc_full);
        __syncwarp();
    } else {
        // =========================== WEIGHT-CHUNK COPIES (the ring reuses the h0 staging area) ===========
        if (lane == 0) {
            const uint64_t w_policy = l2_policy_evict_last();
            mbar_wait(&h0_dead, 0);
            fence_proxy_async_smem();                              // the workers' generic reads of h0 -> async-proxy writes
#pragma unroll 1
            for (int kc = 0; kc < kNumK; ++kc) {
                const int ws = kc & 1;
                if (kc >= 2) mbar_wait(&w_empty[ws], ((kc >> 1) - 1) & 1);
                mbar_arrive_expect_tx(&w_full[ws], 2 * kBTerm);
                bulk_g2s_hint(wring + ws * (2 * kBTerm), a.w1_img + (size_t) kc * kWChunkBytes, 2 * kBTerm, &w_full[ws], w_policy);
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) TP(13);
    if (warp == kVfeWorkerWarps) tmem_dealloc<256>(tmem);
}
