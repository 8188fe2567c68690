// Fused FFN of one encoder layer (included by attention_split.cu inside its anonymous namespace):
//   y = LN-chain( gelu(x W1^T + b1) W2^T + b2 )        x [rows, 192], W1 [384, 192], W2 [192, 384]
// (fullyConnected_gelu_fullyConnected + the addElementWise / LayerNormPlugin pairs behind it, src/dsvt-ai-trt.cpp:494-529,
// :685-697, :750-756) as ONE kernel: the 384-wide hidden rows never leave the SM.  Two kernels (FC + GELU, then FC + norms)
// write 47 MB of hidden rows per layer and read them back through the second kernel's FP32 -> FP16 hi/lo producers; here the
// GELU'd accumulators of the first GEMM are split in registers and written back to TENSOR MEMORY, where the second GEMM
// reads them as its A operand (tcgen05.mma with A in TMEM), 64 hidden columns (one "piece") at a time:
//
//   workers (16 warps)  x tile -> resident FP16 hi/lo image (6 chunks, 96 KB) | per piece: ACC1 -> +b1, GELU, split -> A2 | LN chain
//   issuer  (1 thread)  G1(0) G1(1) G2(0) G1(2) G2(1) ... G1(5) G2(4) G2(5)      G1(p): ACC1[p&1] = X W1_p^T   (N = 64, K = 192)
//                                                                               G2(p): ACC2 += A2[p&1] W2_p^T (N = 192, K = 64)
//   copier  (1 thread)  24 KB weight slots in the issuer's order through a 5-slot ring (W1 piece = 2 slots, W2 piece = 2 slots)
//
// TMEM (512 columns): ACC2 0..191 | ACC1[b] 192+64b.. (FP32) | A2[b] 320+64b..: 32 packed hi columns, 32 packed lo columns.
// The first GEMM runs one piece ahead of the second, so the tensor pipe works on G1(p+1) while the workers turn ACC1(p)
// into A2(p).  Products, their order and the epilogue arithmetic are those of the two-kernel form.
//
// OUT_PROJ = true is the LAYER TAIL (dsvt_attention_tail_ffn_launch): the tile first stages the rows `o` of the attention's
// per-set core, G_o puts the out-projection into ACC2, the shared LayerNorm epilogue drains it as src = norm1(. + x) (tile in
// the dead o image; the weight ring sits behind that tile and keeps streaming), and the FFN's input image is staged from the
// src rows just written -- out-projection, norm1, FFN and the norms behind it are one kernel.
constexpr int kFP = 64;                         // hidden columns per piece
constexpr int kFPieces = 2 * kC / kFP;          // 6
constexpr int kFWSlots = 5;
constexpr int kFSlot = 2 * kBTerm;              // 24576 B
constexpr int kFW1Chunk = 2 * kFP * kBK * 2;    // 8192 B: one K chunk of a W1 piece, [hi | lo][c16][n 64][16 B]
constexpr int kFW1Piece = kNumK * kFW1Chunk;    // 49152 B
constexpr int kFA1 = kNumK * 2 * kATerm;        // 98304 B: the resident x image
constexpr int kFRing = kBM * kLnStride * 4;     // 100352 B: the weight ring starts behind the LayerNorm tile that aliases the image
                                                // (norm1 of the layer-tail form runs while the ring streams the FFN's first pieces)
static_assert(kFRing >= kFA1 && kFRing % 128 == 0, "ring base");
constexpr int kFSmemTotal = kFRing + kFWSlots * kFSlot;                              // 223232 B
constexpr uint32_t kFAcc1 = 192, kFA2 = 320;
constexpr int kFWorkerWarps = 16;               // one CTA per SM (196 KB of shared memory): twice the tile kernel's worker warps
constexpr int kFWorkers = kFWorkerWarps * 32;
constexpr int kFThreads = kFWorkers + 64;       // + warp 16: MMA issue, warp 17: weight-slot copies

struct FfnArgs {
    GemmRole g;               // a0 = x (lda 192), wimg = W2 image (12 chunks of 24 KB), bias = b2, out_mul, LN chain, out
    const uint8_t* w1_img;    // 6 pieces x 48 KB
    const float* bias1;       // [384]
    float out_mul1;
    // OUT_PROJ form (the tile starts with the attention's out-projection + norm1; g.a0 is not read, g.ln_res[0] = src_out):
    const float* o;           // [max_pillars, 192] rows of the per-set core (voxel order)
    const uint8_t* wo_img;    // out-projection: 6 chunk images of 24 KB
    const float* bias_o;      // [192]
    float out_mul_o;
    const int* cover;         // voxel -> (set, token) map of the attention plan: negative = the voxel is in no set (row = 0)
    const float* res1;        // [max_pillars, 192] the layer input x: src = norm1(attention + x)
    const float* gamma1;
    const float* beta1;
    float eps1;
    float* src_out;           // [max_pillars, 192] src rows (read back for the FFN's input image and as the residual of norm2)
};

template <bool OUT_PROJ>
__global__ void __launch_bounds__(kFThreads, 1)
ffn_fused_kernel(const __grid_constant__ FfnArgs args, const int* __restrict__ voxel_num, int max_pillars, int zero_tails)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t a1_full[kNumK], w_full[kFWSlots], w_empty[kFWSlots], acc1_full[2], acc1_empty[2], a2_full[2],
        a2_empty[2], acc2_full, acc_o_full;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias1[2 * kC];

    const GemmRole& g = args.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.z;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const int row_base = tile * kBM;
    float* out = g.out + (size_t) b * max_pillars * g.ld_out;
    if (row_base >= V) {
        if (zero_tails)
            for (int i = tid; i < kBM * (kBN / 4); i += kFThreads) {
                const int rloc = i / (kBN / 4), cc4 = i - rloc * (kBN / 4);
                if (row_base + rloc < max_pillars)
                    stg_zero4(reinterpret_cast<float4*>(out + (size_t) (row_base + rloc) * g.ld_out + cc4 * 4));
            }
        return;
    }
    const float* x = (OUT_PROJ ? args.o : g.a0) + (size_t) b * max_pillars * kC;     // the rows staged first
    if (tid == 0) TP(0);

    if (tid == 0) {
        for (int s = 0; s < kNumK; ++s) mbar_init(&a1_full[s], kFWorkers);
        for (int s = 0; s < kFWSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], kFWorkers);
            mbar_init(&a2_full[s], kFWorkers); mbar_init(&a2_empty[s], 1);
        }
        mbar_init(&acc2_full, 1); mbar_init(&acc_o_full, 1);
        fence_barrier_init();
    }
    if (warp == kFWorkerWarps) tmem_alloc<512>(&tmem_slot);
    for (int t = tid; t < 2 * kC; t += kFThreads) s_bias1[t] = __ldg(args.bias1 + t);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) TP(1);

    if (warp < kFWorkerWarps) {
        // =========================== rows -> resident A image ================================================
        // step = K chunk: row warp*8 + (lane & 7), 16-byte K piece c16 = lane >> 3.  `coherent`: the rows were written by this
        // kernel (src of the OUT_PROJ form) -- read through L2, not through the non-coherent read-only path
        auto stage = [&](const float* rows, bool coherent) {
            constexpr int kDepth = 3;
            const int rl = warp * 8 + (lane & 7), c16 = lane >> 3, row = row_base + rl;
            float buf[kDepth][8];
            auto issue = [&](int kc, float (&d)[8]) {
                const float* p = rows + (size_t) row * kC + kc * kBK + c16 * 8;
                if (row >= V) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[e] = 0.f;
                } else if (coherent) {
                    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p));
                    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p + 4));
                } else {
                    ldg256(p, &d[0]);
                }
            };
#pragma unroll
            for (int kc = 0; kc < kDepth - 1; ++kc) issue(kc, buf[kc]);
#pragma unroll
            for (int kc = 0; kc < kNumK; ++kc) {
                if (kc + kDepth - 1 < kNumK) issue(kc + kDepth - 1, buf[(kc + kDepth - 1) % kDepth]);
                float (&d)[8] = buf[kc % kDepth];
                // the two-kernel form adds the (absent) second operand: v = d + 0
                const float v[8] = {d[0] + 0.f, d[1] + 0.f, d[2] + 0.f, d[3] + 0.f, d[4] + 0.f, d[5] + 0.f, d[6] + 0.f, d[7] + 0.f};
                const uint4 hi = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                const float2 h0 = unpack_h2(hi.x), h1 = unpack_h2(hi.y), h2 = unpack_h2(hi.z), h3 = unpack_h2(hi.w);
                const uint4 lo = make_uint4(pack_h2(v[0] - h0.x, v[1] - h0.y), pack_h2(v[2] - h1.x, v[3] - h1.y),
                                            pack_h2(v[4] - h2.x, v[5] - h2.y), pack_h2(v[6] - h3.x, v[7] - h3.y));
                uint8_t* chunk = smem + kc * (2 * kATerm);
                *reinterpret_cast<uint4*>(chunk + c16 * (kBM * 16) + rl * 16) = hi;
                *reinterpret_cast<uint4*>(chunk + kATerm + c16 * (kBM * 16) + rl * 16) = lo;
                fence_proxy_async_smem();
                mbar_arrive(&a1_full[kc]);
            }
        };
        stage(x, false);
        if (tid == 0) TP(2);
        if (OUT_PROJ) {
            // =========================== out-projection drained through norm1 into the src rows ==================
            // src = LayerNorm(acc * out_mul + bias + x) (src/dsvt-ai-trt.cpp:669-676): the shared LayerNorm-chain epilogue with one
            // stage, its FP32 tile in the (dead) o image -- the weight ring keeps streaming the FFN's first pieces meanwhile.  The
            // FFN's input image is then staged from the src rows just written (L2 hits).
            mbar_wait(&acc_o_full, 0);
            tc_fence_after_sync();
            if (tid == 0) TP(26);
            GemmRole n1 = g;
            n1.bias = args.bias_o; n1.out_mul = args.out_mul_o; n1.cover = args.cover; n1.cover_stride = 0;
            n1.n_ln = 1; n1.ln_eps = args.eps1; n1.ln_res0_written_here = 0;
            n1.ln_res[0] = args.res1; n1.ln_gamma[0] = args.gamma1; n1.ln_beta[0] = args.beta1;
            n1.out = args.src_out;
            ln_chain_epilogue<kFWorkerWarps>(n1, reinterpret_cast<float*>(smem), tmem, warp, lane, tid, row_base, V, b, max_pillars,
                                             args.src_out + (size_t) b * max_pillars * kC, 0);
            tc_fence_before_sync();
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"n"(kFWorkers) : "memory");      // every src row of the tile is visible to the CTA
            stage(args.src_out + (size_t) b * max_pillars * kC, true);        // second phase of a1_full: the src image
            if (tid == 0) TP(27);
        }
        // =========================== per piece: ACC1 -> + b1, GELU, hi/lo split -> A2 (tensor memory) ========
        const int q4 = warp & 3, cq = warp >> 2;               // TMEM lane quarter, 16-column block of the piece
        const uint32_t lane_base = tmem + ((uint32_t) (q4 * 32) << 16);
#pragma unroll 1
        for (int p = 0; p < kFPieces; ++p) {
            const int bb = p & 1;
            mbar_wait(&acc1_full[bb], (p >> 1) & 1);
            tc_fence_after_sync();
            uint32_t r[16];
            tmem_ld16(lane_base + kFAcc1 + bb * kFP + cq * 16, r);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(&acc1_empty[bb]);                      // G1(p + 2) may overwrite the accumulators
            const float* b1 = s_bias1 + p * kFP + cq * 16;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // FC epilogue (acc * out_mul + bias, x 1) and GELU exactly as the tile kernel's, then the producers' split
                const float y0 = gelu_tanh((__uint_as_float(r[2 * c]) * args.out_mul1 + b1[2 * c]) * 1.0f) + 0.f;
                const float y1 = gelu_tanh((__uint_as_float(r[2 * c + 1]) * args.out_mul1 + b1[2 * c + 1]) * 1.0f) + 0.f;
                hi[c] = pack_h2(y0, y1);
                const float2 h = unpack_h2(hi[c]);
                lo[c] = pack_h2(y0 - h.x, y1 - h.y);
            }
            if (p >= 2) { mbar_wait(&a2_empty[bb], ((p >> 1) - 1) & 1); tc_fence_after_sync(); }   // G2(p - 2) has read A2[bb]
            tmem_st8(lane_base + kFA2 + bb * kFP + cq * 8, hi);
            tmem_st8(lane_base + kFA2 + bb * kFP + 32 + cq * 8, lo);
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(&a2_full[bb]);
            if (tid == 0) TP(3 + p);
        }
        // =========================== LayerNorm-chain epilogue ==============================================
        mbar_wait(&acc2_full, 0);                              // every MMA has completed: shared memory is free
        tc_fence_after_sync();
        if (tid == 0) TP(9);
        ln_chain_epilogue<kFWorkerWarps>(g, reinterpret_cast<float*>(smem), tmem, warp, lane, tid, row_base, V, b, max_pillars, out, zero_tails);
    } else if (warp == kFWorkerWarps) {
        // =========================== MMA ISSUE ===========================================================
        {                                                      // the whole warp, converged: see umma_f16_w
            const uint32_t idesc1 = make_idesc(kFmtF16, kBM, kFP), idesc2 = make_idesc(kFmtF16, kBM, kBN);
            const uint32_t sbase = smem_u32(smem), wbase = sbase + kFRing;
            // descriptors = base + (byte offset >> 4): the issuing thread's own instruction stream is on the critical path
            // (36 + 12 MMAs per piece), so nothing but 64-bit adds of compile-time constants stays inside the loops
            const uint64_t a_base = make_smem_desc(sbase, kBM * 16, 128);
            const uint64_t w1_base = make_smem_desc(wbase, kFP * 16, 128), w2_base = make_smem_desc(wbase, kBN * 16, 128);
            int L = 0;                                         // weight-slot loads consumed so far
            if (OUT_PROJ) {                                    // G_o: ACC2 = O Wo^T (drained by norm1 before G2(0) overwrites it)
                const uint64_t wo_base = make_smem_desc(wbase, kBN * 16, 128);
#pragma unroll 1
                for (int kc = 0; kc < kNumK; ++kc, ++L) {
                    const int slot = L % kFWSlots;
                    mbar_wait(&a1_full[kc], 0); __syncwarp();
                    mbar_wait(&w_full[slot], (L / kFWSlots) & 1); __syncwarp();
                    tc_fence_after_sync();
                    const uint64_t ad = a_base + (uint64_t) ((kc * 2 * kATerm) >> 4), wd = wo_base + (uint64_t) ((slot * kFSlot) >> 4);
#pragma unroll
                    for (int ks = 0; ks < kBK / 16; ++ks) {
                        const uint64_t a_hi = ad + (uint64_t) ((ks * 2 * (kBM * 16)) >> 4), a_lo = a_hi + (uint64_t) (kATerm >> 4);
                        const uint64_t b_hi = wd + (uint64_t) ((ks * 2 * (kBN * 16)) >> 4), b_lo = b_hi + (uint64_t) (kBTerm >> 4);
                        umma_f16_w(tmem, a_lo, b_hi, idesc2, (kc | ks) != 0);
                        umma_f16_w(tmem, a_hi, b_lo, idesc2, 1);
                        umma_f16_w(tmem, a_hi, b_hi, idesc2, 1);
                    }
                    umma_commit_w(&w_empty[slot]);
                }
                umma_commit_w(&acc_o_full);
            }
#pragma unroll 1
            for (int i = 0; i < 2 * kFPieces; ++i) {
                const bool g1 = i == 0 || (i != 2 * kFPieces - 1 && (i & 1));
                const int p = i == 0 ? 0 : (i == 2 * kFPieces - 1 ? kFPieces - 1 : (g1 ? (i + 1) >> 1 : (i >> 1) - 1));
                const int bb = p & 1;
                if (lane == 0) TP(14 + i);
                if (g1) {
                    if (p >= 2) { mbar_wait(&acc1_empty[bb], ((p >> 1) - 1) & 1); __syncwarp(); }
                    tc_fence_after_sync();
                    const uint32_t d = tmem + kFAcc1 + bb * kFP;
#pragma unroll
                    for (int half = 0; half < 2; ++half, ++L) {
                        const int slot = L % kFWSlots;
                        mbar_wait(&w_full[slot], (L / kFWSlots) & 1); __syncwarp();
                        tc_fence_after_sync();
                        const uint64_t wd = w1_base + (uint64_t) ((slot * kFSlot) >> 4);
#pragma unroll
                        for (int kcl = 0; kcl < 3; ++kcl) {
                            const int kc = half * 3 + kcl;
                            if (p == 0) { mbar_wait(&a1_full[kc], OUT_PROJ ? 1 : 0); __syncwarp(); tc_fence_after_sync(); }
#pragma unroll
                            for (int ks = 0; ks < kBK / 16; ++ks) {
                                const uint64_t a_hi = a_base + (uint64_t) ((kc * 2 * kATerm + ks * 2 * (kBM * 16)) >> 4);
                                const uint64_t a_lo = a_hi + (uint64_t) (kATerm >> 4);
                                const uint64_t b_hi = wd + (uint64_t) ((kcl * kFW1Chunk + ks * 2 * (kFP * 16)) >> 4);
                                const uint64_t b_lo = b_hi + (uint64_t) ((kFW1Chunk / 2) >> 4);
                                umma_f16_w(d, a_lo, b_hi, idesc1, (kc | ks) != 0);
                                umma_f16_w(d, a_hi, b_lo, idesc1, 1);
                                umma_f16_w(d, a_hi, b_hi, idesc1, 1);
                            }
                        }
                        umma_commit_w(&w_empty[slot]);
                    }
                    umma_commit_w(&acc1_full[bb]);
                } else {
                    mbar_wait(&a2_full[bb], (p >> 1) & 1); __syncwarp();
                    tc_fence_after_sync();
                    const uint32_t a2 = tmem + kFA2 + bb * kFP;
#pragma unroll
                    for (int half = 0; half < 2; ++half, ++L) {
                        const int slot = L % kFWSlots;
                        mbar_wait(&w_full[slot], (L / kFWSlots) & 1); __syncwarp();
                        tc_fence_after_sync();
                        const uint64_t wd = w2_base + (uint64_t) ((slot * kFSlot) >> 4);
#pragma unroll
                        for (int ks = 0; ks < kBK / 16; ++ks) {
                            const uint32_t a_hi = a2 + (half * 2 + ks) * 8, a_lo = a_hi + 32;
                            const uint64_t b_hi = wd + (uint64_t) ((ks * 2 * (kBN * 16)) >> 4);
                            const uint64_t b_lo = b_hi + (uint64_t) (kBTerm >> 4);
                            umma_f16_ts_w(tmem, a_lo, b_hi, idesc2, (p | half | ks) != 0);
                            umma_f16_ts_w(tmem, a_hi, b_lo, idesc2, 1);
                            umma_f16_ts_w(tmem, a_hi, b_hi, idesc2, 1);
                        }
                        umma_commit_w(&w_empty[slot]);
                    }
                    umma_commit_w(&a2_empty[bb]);
                }
            }
            umma_commit_w(&acc2_full);
        }
        __syncwarp();
    } else {
        // =========================== WEIGHT-SLOT COPIES (the issuer's order) =============================
        if (lane == 0) {
            const uint64_t w_policy = l2_policy_evict_last();
            const int nrows = V - row_base < kBM ? V - row_base : kBM;
            const uint32_t bytes = (uint32_t) (nrows * kC * sizeof(float));
            l2_prefetch(x + (size_t) row_base * kC, bytes);
            for (int st = 0; st < g.n_ln; ++st)
                if (g.ln_res[st] != nullptr) l2_prefetch(g.ln_res[st] + ((size_t) b * max_pillars + row_base) * kC, bytes);
            int L0 = 0;
            if (OUT_PROJ) {
                if (args.res1) l2_prefetch(args.res1 + (size_t) row_base * kC, bytes);
#pragma unroll 1
                for (; L0 < kNumK; ++L0) {                     // the out-projection's six chunks, then the FFN's slots
                    const int slot = L0 % kFWSlots;
                    if (L0 >= kFWSlots) mbar_wait(&w_empty[slot], ((L0 / kFWSlots) - 1) & 1);
                    mbar_arrive_expect_tx(&w_full[slot], kFSlot);
                    bulk_g2s_hint(smem + kFRing + slot * kFSlot, args.wo_img + (size_t) L0 * kWChunkBytes, kFSlot, &w_full[slot], w_policy);
                }
            }
#pragma unroll 1
            for (int Lf = 0; Lf < 4 * kFPieces; ++Lf) {
                const int L = L0 + Lf;
                const int i = Lf >> 1, half = Lf & 1, slot = L % kFWSlots;
                const bool g1 = i == 0 || (i != 2 * kFPieces - 1 && (i & 1));
                const int p = i == 0 ? 0 : (i == 2 * kFPieces - 1 ? kFPieces - 1 : (g1 ? (i + 1) >> 1 : (i >> 1) - 1));
                const uint8_t* src = g1 ? args.w1_img + (size_t) p * kFW1Piece + (size_t) half * kFSlot
                                        : g.wimg + (size_t) (2 * p + half) * kWChunkBytes;
                if (L >= kFWSlots) mbar_wait(&w_empty[slot], ((L / kFWSlots) - 1) & 1);
                mbar_arrive_expect_tx(&w_full[slot], kFSlot);
                bulk_g2s_hint(smem + kFRing + slot * kFSlot, src, kFSlot, &w_full[slot], w_policy);
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) TP(13);
    if (warp == kFWorkerWarps) tmem_dealloc<512>(tmem);
}
