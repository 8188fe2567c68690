// Up to eight position-embedding MLPs (Linear(2 -> 192) + BatchNorm + ReLU -> Linear(192 -> 192), src/dsvt-ai-trt.cpp:461-492,
// :603-637) as ONE CTA per 128-row tile with the MLPs as roles in sequence (included by attention_split.cu inside its anonymous
// namespace; same structure as qkv_fused.cuh):
//
//   workers (16 warps)  generate A(0) | generate A(1), drain 0 | generate A(2), drain 1 | ... | drain n-1
//   issuer  (1 warp)    G(0) -> ACC0    G(1) -> ACC1             G(2) -> ACC0
//   copier  (1 thread)  n x 6 weight chunk images through a 3-slot ring
//
// The hidden rows relu(bn(xy W1^T)) are generated from the two window coordinates straight into the GEMM's FP16 hi / lo image
// (they never exist in memory); the drains are the QKV kernel's coalesced epilogue with rows in voxel order.  As the roles of
// the tile GEMM the eight MLPs are 1 928 CTAs with their own set-up, weight-stream start and 9 k-cycle drain each; here a
// role costs a generate (2 k cycles) and a drain, the tensor pipe works underneath.  Arithmetic = the tile GEMM's
// (small_linear_kernel's FMA chain for the hidden rows): outputs are bit-identical.
constexpr int kPWSlots = 3;
constexpr int kPMaxRoles = 8;
constexpr int kPGen = kPMaxRoles * kC * 16;                                    // 24576 B: per role and column (w0, w1, scale, shift)
constexpr int kPSmem = kQA + kPWSlots * 2 * kBTerm + 16 * kQScratch + kPGen;   // 229376 B

struct PosRole {
    const float* x2;          // [max_rows, 2] window coordinates
    const float* gen_blob;    // dsvt_small_linear blob: W1 [192][2] | scale [192] | shift [192]
    const uint8_t* wimg;      // second layer: 6 chunk images of 24 KB
    const float* bias;        // [192]
    float* out;               // [max_rows, 192]
    float out_mul;
};
struct PosArgs { PosRole r[kPMaxRoles]; int n; };

__global__ void __launch_bounds__(kQThreads, 1)
pos_fused_kernel(const __grid_constant__ PosArgs a, const int* __restrict__ voxel_num, int max_rows, int zero_tails)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[kNumK], a_free, w_full[kPWSlots], w_empty[kPWSlots], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, n = a.n;
    int V = voxel_num[0];
    V = V < max_rows ? V : max_rows;
    const int row_base = tile * kBM;
    if (row_base >= V) {                                    // a tile of tail rows: zero-filled on request
        if (zero_tails)
            for (int r = 0; r < n; ++r)
                for (int i = tid; i < kBM * (kBN / 4); i += kQThreads) {
                    const int rloc = i / (kBN / 4), cc4 = i - rloc * (kBN / 4);
                    if (row_base + rloc < max_rows) stg_zero4(reinterpret_cast<float4*>(a.r[r].out + (size_t) (row_base + rloc) * kC + cc4 * 4));
                }
        return;
    }
    float4* s_gen = reinterpret_cast<float4*>(smem + kQA + kPWSlots * 2 * kBTerm + 16 * kQScratch);

    if (tid == 0) {
        for (int s = 0; s < kNumK; ++s) mbar_init(&a_full[s], kQWorkers);
        for (int s = 0; s < kPWSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(&a_free, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kQWorkers); }
        fence_barrier_init();
    }
    if (warp == kQWorkerWarps) tmem_alloc<512>(&tmem_slot);
    for (int t = tid; t < n * kC; t += kQThreads) {
        const int r = t / kC, c = t - r * kC;
        const float* gb = a.r[r].gen_blob;
        s_gen[t] = make_float4(__ldg(gb + 2 * c), __ldg(gb + 2 * c + 1), __ldg(gb + 2 * kC + c), __ldg(gb + 3 * kC + c));
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp < kQWorkerWarps) {
        const int rl = warp * 8 + (lane & 7), c16 = lane >> 3, srow = row_base + rl;      // generate: row, 16-byte K piece
        auto generate = [&](int role) {
            float2 xy = make_float2(0.f, 0.f);
            if (srow < V) xy = __ldg(reinterpret_cast<const float2*>(a.r[role].x2) + srow);
            const float4* gw = s_gen + role * kC;
#pragma unroll
            for (int kc = 0; kc < kNumK; ++kc) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {              // same arithmetic as small_linear_kernel<2>: FC, Scale (folded BatchNorm), ReLU
                    const float4 wv = gw[kc * kBK + c16 * 8 + e];
                    float acc = fmaf(xy.x, wv.x, 0.f);
                    acc = fmaf(xy.y, wv.y, acc);
                    acc = fmaf(acc, wv.z, wv.w);
                    v[e] = (srow < V ? fmaxf(acc, 0.f) : 0.f) + 0.f;      // (+ 0: the tile GEMM's producers add the absent second operand)
                }
                const uint4 hi = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                const float2 h0 = unpack_h2(hi.x), h1 = unpack_h2(hi.y), h2 = unpack_h2(hi.z), h3 = unpack_h2(hi.w);
                const uint4 lo = make_uint4(pack_h2(v[0] - h0.x, v[1] - h0.y), pack_h2(v[2] - h1.x, v[3] - h1.y),
                                            pack_h2(v[4] - h2.x, v[5] - h2.y), pack_h2(v[6] - h3.x, v[7] - h3.y));
                uint8_t* chunk = smem + kc * (2 * kATerm);
                *reinterpret_cast<uint4*>(chunk + c16 * (kBM * 16) + rl * 16) = hi;
                *reinterpret_cast<uint4*>(chunk + kATerm + c16 * (kBM * 16) + rl * 16) = lo;
                fence_proxy_async_smem();
                mbar_arrive(&a_full[kc]);
            }
        };
        // drain: TMEM lane quarter q4, 48-column block cb, through the warp's swizzled scratch (see qkv_fused.cuh)
        const int q4 = warp & 3, cb = warp >> 2;
        int srow_out[4];                                    // rows this lane stores: (lane >> 2) + 8 i of the warp's 32
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int g = row_base + q4 * 32 + (lane >> 2) + 8 * i;
            srow_out[i] = g < V ? g : ((zero_tails && g < max_rows) ? -2 - g : -1);       // -1: not written, <= -2: zero row -2 - g
        }
        float4* scr = reinterpret_cast<float4*>(smem + kQA + kPWSlots * 2 * kBTerm + warp * kQScratch);
        auto drain = [&](int role) {
            const int bb = role & 1;
            mbar_wait(&acc_full[bb], (role >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t tl = tmem + bb * 192 + ((uint32_t) (q4 * 32) << 16) + cb * 48;
            const float om = a.r[role].out_mul;
            float* base = a.r[role].out;
#pragma unroll
            for (int j0 = 0; j0 < 48; j0 += 16) {
                uint32_t r[16];
                tmem_ld16(tl + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    scr[lane * 4 + (j ^ (lane & 3))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                __syncwarp();
                const int c4 = lane & 3, col = cb * 48 + j0 + 4 * c4;
                const float4 bv = __ldg(reinterpret_cast<const float4*>(a.r[role].bias + col));
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int sr = (lane >> 2) + 8 * i;
                    const float4 v = scr[sr * 4 + (c4 ^ (sr & 3))];
                    const float4 ov = make_float4((v.x * om + bv.x) * 1.0f, (v.y * om + bv.y) * 1.0f, (v.z * om + bv.z) * 1.0f, (v.w * om + bv.w) * 1.0f);
                    if (srow_out[i] >= 0) *reinterpret_cast<float4*>(base + (size_t) srow_out[i] * kC + col) = ov;
                    else if (srow_out[i] <= -2) stg_zero4(reinterpret_cast<float4*>(base + (size_t) (-2 - srow_out[i]) * kC + col));
                }
                __syncwarp();
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[bb]);
        };
#pragma unroll 1
        for (int role = 0; role < n; ++role) {
            if (role >= 1) mbar_wait(&a_free, (role - 1) & 1);      // G(role - 1) has read the image
            generate(role);
            if (role >= 1) drain(role - 1);
        }
        drain(n - 1);
    } else if (warp == kQWorkerWarps) {
        // =========================== MMA ISSUE (converged warp) ==========================================
        const uint32_t idesc = make_idesc(kFmtF16, kBM, kBN);
        const uint32_t sbase = smem_u32(smem), wbase = sbase + kQA;
        const uint64_t a_base = make_smem_desc(sbase, kBM * 16, 128), w_base = make_smem_desc(wbase, kBN * 16, 128);
        int L = 0;
#pragma unroll 1
        for (int role = 0; role < n; ++role) {
            const uint32_t d = tmem + (role & 1) * 192;
            if (role >= 2) { mbar_wait(&acc_empty[role & 1], ((role >> 1) - 1) & 1); __syncwarp(); tc_fence_after_sync(); }
#pragma unroll 1
            for (int kc = 0; kc < kNumK; ++kc, ++L) {
                mbar_wait(&a_full[kc], role & 1); __syncwarp();
                const int slot = L % kPWSlots;
                mbar_wait(&w_full[slot], (L / kPWSlots) & 1); __syncwarp();
                tc_fence_after_sync();
                const uint64_t ad = a_base + (uint64_t) ((kc * 2 * kATerm) >> 4), wd = w_base + (uint64_t) ((slot * 2 * kBTerm) >> 4);
#pragma unroll
                for (int ks = 0; ks < kBK / 16; ++ks) {
                    const uint64_t a_hi = ad + (uint64_t) ((ks * 2 * (kBM * 16)) >> 4), a_lo = a_hi + (uint64_t) (kATerm >> 4);
                    const uint64_t b_hi = wd + (uint64_t) ((ks * 2 * (kBN * 16)) >> 4), b_lo = b_hi + (uint64_t) (kBTerm >> 4);
                    umma_f16_w(d, a_lo, b_hi, idesc, (kc | ks) != 0);
                    umma_f16_w(d, a_hi, b_lo, idesc, 1);
                    umma_f16_w(d, a_hi, b_hi, idesc, 1);
                }
                umma_commit_w(&w_empty[slot]);
            }
            umma_commit_w(&acc_full[role & 1]);
            umma_commit_w(&a_free);
        }
        __syncwarp();
    } else {
        // =========================== WEIGHT-CHUNK COPIES ==================================================
        if (lane == 0) {
            const uint64_t w_policy = l2_policy_evict_last();
#pragma unroll 1
            for (int L = 0; L < n * kNumK; ++L) {
                const int slot = L % kPWSlots, role = L / kNumK, kc = L - role * kNumK;
                if (L >= kPWSlots) mbar_wait(&w_empty[slot], ((L / kPWSlots) - 1) & 1);
                mbar_arrive_expect_tx(&w_full[slot], 2 * kBTerm);
                bulk_g2s_hint(smem + kQA + slot * (2 * kBTerm), a.r[role].wimg + (size_t) kc * kWChunkBytes, 2 * kBTerm, &w_full[slot], w_policy);
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kQWorkerWarps) tmem_dealloc<512>(tmem);
}
