// Q / K / V projection of a set-attention layer as ONE CTA per 128-row tile (included by attention_split.cu inside its anonymous
// namespace): the three roles of the tile GEMM share their rows, so a tile-wide CTA converts `x + pos` ONCE for Q and K (the
// tile kernel converts it in two CTAs), re-stages `x` for V while the K projection runs, and drains one accumulator while the
// tensor pipe fills the other:
//
//   workers (16 warps)  stage(x + pos) | epilogue Q | stage(x) | epilogue K | epilogue V
//   issuer  (1 warp)    G_q -> ACC0      G_k -> ACC1           | G_v -> ACC0
//   copier  (1 thread)  Wq, Wk, Wv chunk images (18 x 24 KB) through a 4-slot ring
//
// One CTA per SM (224 KB: the resident FP16 hi / lo image of the tile, 96 KB, the weight ring and 2 KB of transposition scratch
// per worker warp), 384 TMEM columns.  Epilogue: every tmem_ld16 leaves a lane with 16 consecutive columns of its own row; stored
// from there, each instruction would touch 32 different lines and the LSU's one-line-per-cycle stage becomes the limit (8.5 k
// cycles per role, measured) -- so the 32 x 16 block goes through the warp's swizzled scratch and leaves as 8 rows x 64
// contiguous bytes per instruction (token-order rows from the attention plan; K / V rows in the core kernel's head-padded layout).  Operand images, MMA order and epilogue arithmetic are the tile kernel's: q / k|v rows are bit-identical.
constexpr int kQWSlots = 4;
constexpr int kQA = kNumK * 2 * kATerm;                   // 98304 B
constexpr int kQScratch = 32 * 16 * 4;                    // 2048 B per worker warp: 32 rows x 16 columns, XOR-swizzled float4s
constexpr int kQSmem = kQA + kQWSlots * 2 * kBTerm + 16 * kQScratch;      // 229376 B
constexpr int kQWorkerWarps = 16, kQWorkers = kQWorkerWarps * 32, kQThreads = kQWorkers + 64;

struct QkvArgs {
    const float* x;           // [B, max_pillars, 192]
    const float* pos;         // [B, max_pillars, 192]; or, with pos_cell != nullptr, a TABLE [win_x * win_y, 192] indexed by the
                              // voxel's cell in its window: the position embedding is a function of (cx, cy) only
    const int* pos_cell;      // [B, max_pillars, 3] (cz, cy, cx) = WindowPartitionPlugin output 4 (coors_in_win_2d), or nullptr
    int win_x;
    const uint8_t* wimg;      // roles Q, K, V: 3 x 6 chunk images of 24 KB
    const float* bias;        // [3][192]
    float out_mul[3];
    float q_post_mul;         // 1 / sqrt(head dim), applied after the biased projection like the reference's division
    float* qbuf;              // [B, max_pillars, 192]   rows in token order
    float* kvbuf;             // [B, max_pillars, 392]   K row | V row per token, heads 4-7 sixteen bytes further
    const int* plan;
    size_t plan_stride;
};

__global__ void __launch_bounds__(kQThreads, 1)
qkv_fused_kernel(const __grid_constant__ QkvArgs a, const int* __restrict__ voxel_num, int max_pillars, int max_sets)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t a_full[kNumK], a_free, w_full[kQWSlots], w_empty[kQWSlots], acc_full[2], acc_empty0;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.z;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const int row_base = tile * kBM;
    if (row_base >= V) return;                              // q / k|v rows beyond the valid count are never read
    const float* x = a.x + (size_t) b * max_pillars * kC;
    const bool table = a.pos_cell != nullptr;
    const float* pos = table ? a.pos : a.pos + (size_t) b * max_pillars * kC;
    if (tid == 0) TP(0);

    if (tid == 0) {
        for (int s = 0; s < kNumK; ++s) mbar_init(&a_full[s], kQWorkers);
        for (int s = 0; s < kQWSlots; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(&a_free, 1); mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1); mbar_init(&acc_empty0, kQWorkers);
        fence_barrier_init();
    }
    if (warp == kQWorkerWarps) tmem_alloc<512>(&tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) TP(1);

    if (warp < kQWorkerWarps) {
        const int rl = warp * 8 + (lane & 7), c16 = lane >> 3, srow = row_base + rl;      // staging: row, 16-byte K piece
        size_t prow = (size_t) srow * kC;                   // this row's position embedding: its own row, or the table row of its cell
        if (table && srow < V) {
            const int* cell = a.pos_cell + ((size_t) b * max_pillars + srow) * 3;
            prow = (size_t) (__ldg(cell + 1) * a.win_x + __ldg(cell + 2)) * kC;
        }
        // stage(): the tile's rows x + pos -> FP16 hi / lo chunk images (the x-alone image for V is staged from registers below)
        auto stage = [&](bool with_pos) {
            constexpr int kDepth = 3;
            float buf[kDepth][16];
            auto issue = [&](int kc, float (&d)[16]) {
                if (srow < V) {
                    ldg256(x + (size_t) srow * kC + kc * kBK + c16 * 8, &d[0]);
                    if (with_pos) ldg256(pos + prow + kc * kBK + c16 * 8, &d[8]);
                    else {
#pragma unroll
                        for (int e = 8; e < 16; ++e) d[e] = 0.f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) d[e] = 0.f;
                }
            };
#pragma unroll
            for (int kc = 0; kc < kDepth - 1; ++kc) issue(kc, buf[kc]);
#pragma unroll
            for (int kc = 0; kc < kNumK; ++kc) {
                if (kc + kDepth - 1 < kNumK) issue(kc + kDepth - 1, buf[(kc + kDepth - 1) % kDepth]);
                float (&d)[16] = buf[kc % kDepth];
                const float v[8] = {d[0] + d[8], d[1] + d[9], d[2] + d[10], d[3] + d[11], d[4] + d[12], d[5] + d[13], d[6] + d[14], d[7] + d[15]};
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
        // epilogue: TMEM lane quarter q4, 48-column block cb; this lane's row -> its token position (set-major order)
        const int q4 = warp & 3, cb = warp >> 2;
        const int grow = row_base + q4 * 32 + lane;
        int orow = -1;
        if (grow < V) {
            const PlanView pv = plan_view(const_cast<int*>(a.plan) + (size_t) b * a.plan_stride, max_sets, max_pillars);
            const int su = __ldg(pv.vox_su + grow);
            if (su >= 0) {
                const int t = __ldg(pv.set_off + (su >> 6)) + (su & 63);
                if (t < max_pillars) orow = t;
            }
        }
        // rows this lane STORES in the read phase of the transposition: row sr = (lane >> 2) + 8 i, float4 (lane & 3) of it
        int srow_out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) srow_out[i] = __shfl_sync(0xffffffffu, orow, (lane >> 2) + 8 * i);
        float4* scr = reinterpret_cast<float4*>(smem + kQA + kQWSlots * 2 * kBTerm + warp * kQScratch);
        auto epilogue = [&](int role, uint32_t acc) {
            const uint32_t tl = tmem + acc + ((uint32_t) (q4 * 32) << 16) + cb * 48;
            const float om = a.out_mul[role], pm = role == 0 ? a.q_post_mul : 1.0f;
            const size_t ld = role == 0 ? kC : kKvTok;
            float* base = (role == 0 ? a.qbuf : a.kvbuf + (role == 2 ? kKvRow : 0)) + (size_t) b * max_pillars * ld;
#pragma unroll
            for (int j0 = 0; j0 < 48; j0 += 16) {
                uint32_t r[16];
                tmem_ld16(tl + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)                  // float4 j of row `lane` lands in slot j ^ (lane & 3)
                    scr[lane * 4 + (j ^ (lane & 3))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                __syncwarp();
                const int c4 = lane & 3, col = cb * 48 + j0 + 4 * c4;
                const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + role * kC + col));
                const int shift = role != 0 && col >= 96 ? 4 : 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {                // 8 rows x 64 contiguous bytes per store instruction
                    const int sr = (lane >> 2) + 8 * i;
                    const float4 v = scr[sr * 4 + (c4 ^ (sr & 3))];
                    // out_mul is a power of two: the product is exact, so this is one rounding of (acc + bias)
                    const float4 ov = make_float4((v.x * om + bb.x) * pm, (v.y * om + bb.y) * pm, (v.z * om + bb.z) * pm, (v.w * om + bb.w) * pm);
#ifdef DSVT_EXP_NOSTORE    // experiment: the drain without its stores (ov.x can never be this value)
                    if (srow_out[i] >= 0 && ov.x == 123456.789f) *reinterpret_cast<float4*>(base + (size_t) srow_out[i] * ld + col + shift) = ov;
#else
                    if (srow_out[i] >= 0) *reinterpret_cast<float4*>(base + (size_t) srow_out[i] * ld + col + shift) = ov;
#endif
                }
                __syncwarp();
            }
        };
        stage(true);
        if (tid == 0) TP(2);
        // the x rows of the second image are requested NOW (48 registers), so that they arrive while Q is drained: staged after
        // the drain from memory, the second image costs a full load round trip (5.8 k cycles of a 43 k-cycle CTA, measured)
        float xr[kNumK][8];
#pragma unroll
        for (int kc = 0; kc < kNumK; ++kc) {
            if (srow < V) ldg256(x + (size_t) srow * kC + kc * kBK + c16 * 8, xr[kc]);
            else {
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[kc][e] = 0.f;
            }
        }
        mbar_wait(&acc_full[0], 0);
        tc_fence_after_sync();
        epilogue(0, 0);
        tc_fence_before_sync();
        mbar_arrive(&acc_empty0);                             // G_v may overwrite ACC0
        if (tid == 0) TP(3);
        mbar_wait(&a_free, 0);                                // Q and K have read the (x + pos) image
#pragma unroll
        for (int kc = 0; kc < kNumK; ++kc) {
            float (&d)[8] = xr[kc];
            const float v[8] = {d[0] + 0.f, d[1] + 0.f, d[2] + 0.f, d[3] + 0.f, d[4] + 0.f, d[5] + 0.f, d[6] + 0.f, d[7] + 0.f};
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
        if (tid == 0) TP(4);
        mbar_wait(&acc_full[1], 0);
        tc_fence_after_sync();
        epilogue(1, 192);
        if (tid == 0) TP(5);
        mbar_wait(&acc_full[0], 1);
        tc_fence_after_sync();
        epilogue(2, 0);
        if (tid == 0) TP(6);
    } else if (warp == kQWorkerWarps) {
        // =========================== MMA ISSUE (converged warp) ==========================================
        const uint32_t idesc = make_idesc(kFmtF16, kBM, kBN);
        const uint32_t sbase = smem_u32(smem), wbase = sbase + kQA;
        const uint64_t a_base = make_smem_desc(sbase, kBM * 16, 128), w_base = make_smem_desc(wbase, kBN * 16, 128);
        int L = 0;
#pragma unroll 1
        for (int role = 0; role < 3; ++role) {
            const uint32_t d = tmem + (role == 1 ? 192 : 0);
            if (role == 2) { mbar_wait(&acc_empty0, 0); __syncwarp(); tc_fence_after_sync(); }
#pragma unroll 1
            for (int kc = 0; kc < kNumK; ++kc, ++L) {
                if (role != 1) { mbar_wait(&a_full[kc], role == 0 ? 0 : 1); __syncwarp(); }
                const int slot = L % kQWSlots;
                mbar_wait(&w_full[slot], (L / kQWSlots) & 1); __syncwarp();
                tc_fence_after_sync();
                if (lane == 0 && role == 0) TP(14 + kc);
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
            umma_commit_w(&acc_full[role == 1 ? 1 : 0]);
            if (role == 1) umma_commit_w(&a_free);            // Q and K have read the (x + pos) image
        }
        __syncwarp();
    } else {
        // =========================== WEIGHT-CHUNK COPIES ==================================================
        if (lane == 0) {
            const uint64_t w_policy = l2_policy_evict_last();
            const int nrows = V - row_base < kBM ? V - row_base : kBM;
            const uint32_t bytes = (uint32_t) (nrows * kC * sizeof(float));
            l2_prefetch(x + (size_t) row_base * kC, bytes);
            if (!table) l2_prefetch(pos + (size_t) row_base * kC, bytes);
#pragma unroll 1
            for (int L = 0; L < 3 * kNumK; ++L) {
                const int slot = L % kQWSlots;
                if (L >= kQWSlots) mbar_wait(&w_empty[slot], ((L / kQWSlots) - 1) & 1);
                mbar_arrive_expect_tx(&w_full[slot], 2 * kBTerm);
                bulk_g2s_hint(smem + kQA + slot * (2 * kBTerm), a.wimg + (size_t) L * kWChunkBytes, 2 * kBTerm, &w_full[slot], w_policy);
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) TP(13);
    if (warp == kQWorkerWarps) tmem_dealloc<512>(tmem);
}
