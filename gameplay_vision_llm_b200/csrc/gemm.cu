// gemm.cu — K2: persistent, warp-specialised tcgen05 GEMM with a fused epilogue.
//
//   out[M,N] = act(A[M,K] . W[N,K]^T + bias) + residual        (optionally with LayerNorm(A) folded in)
//
// Both operands are K-major bf16, so a [rows x 64] box lands in shared memory as rows of 128 bytes —
// exactly one SWIZZLE_128B atom wide — and the same bytes are addressed by the UMMA shared-memory
// descriptor.  The grid is one 2-CTA cluster per SM pair; a cluster loops over "super tiles" of
// 256 x BN outputs (m-major order: clusters running at the same time share A panels, the whole weight
// stays in the 126 MB L2) and runs ONE tcgen05.mma.cta_group::2 (256 x BN x 16) per k-step: each CTA
// stages only its own 128 A rows and HALF of the W tile (BN/2 rows), the hardware reads the other half
// from the peer's shared memory.  Per CTA and k-block that is 16 KB + BN*64 B of shared-memory traffic
// instead of 16 KB + BN*128 B, which takes the kernel off the operand-feed limit (v1, one CTA per tile:
// 15 TB/s of L2 reads at 64 % tensor-pipe utilisation, profiles/r01_v1_*).
//
//   warp 0       TMA producer (both CTAs): own A box 128x64 + own half W box (BN/2)x64; completion on the
//                LEADER's full barrier
//   warp 1       tcgen05.mma issuer (leader CTA, one elected lane); tcgen05.commit multicasts to both CTAs'
//                empty / tmem_full barriers
//   warps 2..    epilogue, BN/16 warps = BN/64 per TMEM lane quadrant, 64 columns (two 32-column chunks) each:
//                tcgen05.ld 32 lanes x 32 columns -> bias / LayerNorm fold / GELU / residual -> bf16 into a
//                swizzled staging tile -> TMA bulk store (fp32 outputs: direct stores)
//
// The accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.  The epilogue is a single-warp-latency-bound instruction stream (ncu: ~5.5 cycles per
// issued instruction per warp), so (a) it runs on 4 warps per scheduler and (b) the hot flag combinations
// are compiled as separate instantiations (EF_* below) so that no per-element branch survives; any other
// combination takes the generic instantiation with run-time flags.  Ragged M / N / K edges are handled by
// TMA (out-of-bounds reads are zero, including the phantom second half of an odd last super tile, and
// out-of-bounds parts of stored boxes are clipped) and by row / column guards on the direct stores.
#include "common.cuh"

#include <cstdlib>

namespace gvl {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int kMaxLnSlotPairs = 10;  // LayerNorm-fusion consumers read at most 20 partial-sum slots per row

// epilogue flavour, compile-time unless EF_ANY is set
enum : uint32_t { EF_ACT = 3u, EF_RES = 4u, EF_LN = 8u, EF_STATS = 16u, EF_F32 = 32u, EF_ANY = 64u,
                  EF_TMARES = 128u /* residual tile fetched by TMA into the staging tile (compile-time flavours only) */ };

struct GemmParams {
    const float* bias;
    const __nv_bfloat16* residual;
    int ldr;
    int res_row_mod;
    void* out;
    int ldo;
    int out_f32;
    int M, N, K;
    int act;
    int m_pairs, n_tiles, k_blocks;
    int debug;  // tuning experiments only (GVL_GEMM_DEBUG): 1 = stop issuing TMA loads once the ring was filled,
                // 2 = always load the same boxes, 8 = skip the output stores
    // LayerNorm fusion (gvl_gemm_fusion): producer side writes per-slab partial row sums of its bf16 outputs,
    // consumer side normalises the rows of A algebraically in the epilogue
    float* stats_out;       // [M, stats_slots, 2] or null
    int stats_slots;
    const float* ln_stats;  // [M, ln_slots, 2] or null
    const float* ln_c1;     // [N]
    int ln_slots;
    float ln_inv_d, ln_eps;
};

template <int BN>
struct GemmCfg {
    static constexpr int EPI_WARPS = BN / 16;          // BN/64 per TMEM lane quadrant
    static constexpr int THREADS = 64 + EPI_WARPS * 32;
    static constexpr int COLS_PER_WARP = 64;           // two 32-column chunks
    static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;  // per CTA
    static constexpr int STAGES = BN == 256 ? 5 : (BN == 192 ? 6 : 7);
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int BAR_BYTES = (2 * STAGES + 4 + EPI_WARPS) * 8 + 16;  // + one residual-tile barrier per epilogue warp
    // bf16 outputs leave through per-warp staging tiles (32 rows x 64 B, SWIZZLE_64B) and TMA stores
    static constexpr int OUT_TILE_BYTES = 32 * 64;
    static constexpr int OUT_STAGE_BYTES = EPI_WARPS * OUT_TILE_BYTES;
    static constexpr int BIAS_BYTES = 2 * EPI_WARPS * COLS_PER_WARP * 4;  // bias (or c2) and c1, private per warp
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + BIAS_BYTES + BAR_BYTES + 1024;
};

// Residual values of one 32-column chunk of this thread's output row (4 x 8 bf16).  Issued one chunk ahead of
// their use — and, for the first chunk of a tile, before the wait on the accumulator — so that their L2 / HBM
// latency never sits on the epilogue's critical path.
__device__ __forceinline__ void load_residual_chunk(const GemmParams& p, int row, int c0, uint4 (&rv)[4]) {
    const bool row_ok = row < p.M;
    const int rrow = p.res_row_mod > 0 ? (row % p.res_row_mod) : row;
    const __nv_bfloat16* rp = p.residual + (size_t)rrow * p.ldr + c0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        rv[g] = (row_ok && c0 + g * 8 < p.N) ? *reinterpret_cast<const uint4*>(rp + g * 8) : make_uint4(0, 0, 0, 0);
}

// mean * rstd and rstd of one row of A from the producer's partial sums (fixed order: deterministic).  Called
// before the wait on the accumulator so the loads stay off the epilogue's critical path.
__device__ __forceinline__ void load_ln_row(const GemmParams& p, int row, float& ln_rstd, float& ln_mr) {
    ln_rstd = 1.0f;
    ln_mr = 0.0f;
    if (row >= p.M) return;
    if (p.ln_slots == 0) {  // finalised by gvl_ln_finalize: (rstd, mean * rstd) per row
        const float2 f = reinterpret_cast<const float2*>(p.ln_stats)[row];
        ln_rstd = f.x;
        ln_mr = f.y;
        return;
    }
    // slots come in pairs: 16-byte loads, all issued before the first use so the row costs one L2 round trip
    const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + (size_t)row * p.ln_slots * 2);
    const int pairs = p.ln_slots >> 1;
    float4 t[kMaxLnSlotPairs];
#pragma unroll
    for (int i = 0; i < kMaxLnSlotPairs; ++i) t[i] = i < pairs ? sp[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxLnSlotPairs; ++i) {
        s1 += t[i].x;
        s2 += t[i].y;
        s1 += t[i].z;
        s2 += t[i].w;
    }
    const float mean = s1 * p.ln_inv_d;
    const float var = fmaxf(s2 * p.ln_inv_d - mean * mean, 0.0f);
    ln_rstd = rsqrtf(var + p.ln_eps);
    ln_mr = mean * ln_rstd;
}

// Epilogue of one warp's 32 rows x 64 columns of a 128 x BN accumulator tile: TMEM lane quadrant q, columns
// n0 .. n0+63 of the output (t_addr points at the first of them).  rv holds the residual of the first chunk.
template <int BN, uint32_t EF>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, const float* myBias, uint32_t t_addr, int row0, int n0,
                                              int lane, uint4 (&rv)[4], float ln_rstd, float ln_mr,
                                              const CUtensorMap* tmC, uint8_t* stage, bool& store_pending,
                                              const CUtensorMap* tmR, uint64_t* res_bar, uint32_t& res_phase) {
    using Cfg = GemmCfg<BN>;
    constexpr bool ANY = (EF & EF_ANY) != 0;
    // residual through TMA: the 32 x 32 residual tile lands in this warp's staging tile (same SWIZZLE_64B layout the
    // store uses), every thread adds its own 16-byte slots and writes the result back in place — full 64-byte row
    // segments from L2 instead of 32 scattered 16-byte loads per warp instruction, and no residual registers
    constexpr bool TMARES = !ANY && (EF & EF_TMARES) != 0;
    const int act = ANY ? p.act : (int)(EF & EF_ACT);
    const bool has_res = ANY ? p.residual != nullptr : (EF & EF_RES) != 0;
    const bool ln = ANY ? p.ln_stats != nullptr : (EF & EF_LN) != 0;
    const bool stats = ANY ? p.stats_out != nullptr : (EF & EF_STATS) != 0;
    const bool f32 = ANY ? p.out_f32 != 0 : (EF & EF_F32) != 0;
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    // fused LayerNorm of the A rows: out = rstd * (acc - mean * c1[n]) + c2[n]  (c2 arrives through the bias slot)
    const uint32_t bias_s = smem_u32(myBias);
    const uint32_t c1_s = bias_s + Cfg::EPI_WARPS * Cfg::COLS_PER_WARP * 4;
    const uint32_t stage_s = smem_u32(stage);
    float st1 = 0.f, st2 = 0.f;  // partial row sums of this warp's bf16 outputs (stats_out)
#pragma unroll 1
    for (int chunk = 0; chunk < Cfg::COLS_PER_WARP / 32; ++chunk) {
        const int c0 = n0 + chunk * 32;
        if (c0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(t_addr + (uint32_t)(chunk * 32), r);
        uint4 rn[4];  // next chunk's residual, in flight while this chunk is finished
        if (!TMARES && has_res && chunk + 1 < Cfg::COLS_PER_WARP / 32) load_residual_chunk(p, row, c0 + 32, rn);
        if (!f32 && store_pending) {  // the staging tile is reused: the previous store must have read it
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
        }
        if (TMARES) {
            if (lane == 0) {
                mbar_arrive_expect_tx(res_bar, 32 * 64);
                tma_load_2d(stage, tmR, res_bar, c0, row0);
            }
            __syncwarp();
        }
        tmem_ld_wait();
        if (TMARES) {
            mbar_wait(res_bar, res_phase);
            res_phase ^= 1u;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int col = c0 + g * 8;
            const float4 b0 = ld_shared_f4(bias_s + (chunk * 32 + g * 8) * 4);
            const float4 b1 = ld_shared_f4(bias_s + (chunk * 32 + g * 8 + 4) * 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float v[8];
            if (ln) {
                const float4 k0 = ld_shared_f4(c1_s + (chunk * 32 + g * 8) * 4);
                const float4 k1 = ld_shared_f4(c1_s + (chunk * 32 + g * 8 + 4) * 4);
                const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(r[g * 8 + j]), ln_rstd, fmaf(-ln_mr, kk[j], bb[j]));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) + bb[j];
            }
            if (act == GVL_ACT_GELU_TANH) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_tanh_f(v[j]);
            } else if (act == GVL_ACT_GELU_ERF) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_erf_f(v[j]);
            }
            if (TMARES) {
                const float4 rf = ld_shared_f4(stage_s + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4));
                const uint32_t rw[4] = {__float_as_uint(rf.x), __float_as_uint(rf.y), __float_as_uint(rf.z),
                                        __float_as_uint(rf.w)};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[2 * e] += bf16_lo(rw[e]);
                    v[2 * e + 1] += bf16_hi(rw[e]);
                }
            } else if (has_res) {
                v[0] += bf16_lo(rv[g].x); v[1] += bf16_hi(rv[g].x);
                v[2] += bf16_lo(rv[g].y); v[3] += bf16_hi(rv[g].y);
                v[4] += bf16_lo(rv[g].z); v[5] += bf16_hi(rv[g].z);
                v[6] += bf16_lo(rv[g].w); v[7] += bf16_hi(rv[g].w);
            }
            if (f32) {
                if (row_ok && col < p.N) {
                    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col;
                    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            } else {
                uint4 ov;
                ov.x = pack_bf16x2(v[0], v[1]);
                ov.y = pack_bf16x2(v[2], v[3]);
                ov.z = pack_bf16x2(v[4], v[5]);
                ov.w = pack_bf16x2(v[6], v[7]);
                // staging tile row = lane, 16-byte slot g, SWIZZLE_64B: slot ^= (row >> 1) & 3.  Rows / columns beyond
                // M / N are staged too (finite garbage) and clipped by the TMA store.
                st_shared_v4(stage_s + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4), ov);
                if (stats && col < p.N) {  // statistics of the values as stored (bf16-rounded)
                    const uint32_t w4[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = bf16_lo(w4[e]), a1 = bf16_hi(w4[e]);
                        st1 += a0 + a1;
                        st2 = fmaf(a0, a0, fmaf(a1, a1, st2));
                    }
                }
            }
        }
        if (!f32) {
            // hand the staged 32 x 32 tile to the TMA engine
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(p.debug & 8)) {
                tma_store_2d(tmC, stage, c0, row0);
                tma_store_commit();
            }
            store_pending = true;
        }
        if (!TMARES && has_res && chunk + 1 < Cfg::COLS_PER_WARP / 32) {
#pragma unroll
            for (int g = 0; g < 4; ++g) rv[g] = rn[g];
        }
    }
    if (stats && row_ok) {  // column groups entirely beyond N write zeros: every slot of the row is defined
        const int slot = n0 / Cfg::COLS_PER_WARP, real = (BN / Cfg::COLS_PER_WARP) * p.n_tiles;
        float2* so = reinterpret_cast<float2*>(p.stats_out) + (size_t)row * p.stats_slots;
        so[slot] = make_float2(st1, st2);
        // the slot count is padded to an even number for the consumer's 16-byte loads
        if ((real & 1) && slot == real - 1) so[real] = make_float2(0.f, 0.f);
    }
}

template <int BN, uint32_t EF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GemmCfg<BN>::THREADS, 1)
gemm_bf16_cg2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                     const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool ANY = (EF & EF_ANY) != 0;
    constexpr bool TMARES = !ANY && (EF & EF_TMARES) != 0;

    extern __shared__ uint8_t smem_raw[];
    // the dynamic smem window starts at the same CTA-relative offset in both CTAs of the cluster, so the
    // aligned carve-up below is identical in both (required by the peer's smem reads / remote arrives)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint8_t* sOutStage = smem + STAGES * Cfg::STAGE_BYTES;  // [epilogue warp][32 rows x 64 B], 1024-byte aligned
    float* sBias = reinterpret_cast<float*>(sOutStage + Cfg::OUT_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
    uint64_t* full_bar = bars;                       // used in the leader only
    uint64_t* empty_bar = bars + STAGES;             // per CTA: "this stage may be overwritten"
    uint64_t* tmem_full_bar = bars + 2 * STAGES;     // per CTA: accumulator stage ready
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // leader only: both CTAs' epilogues drained the stage
    uint64_t* res_bar = bars + 2 * STAGES + 4;       // per epilogue warp: residual tile landed (EF_TMARES)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + Cfg::EPI_WARPS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    pdl_trigger();  // the next kernel's CTAs may take this SM as soon as this CTA leaves it

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);   // leader producer's arrive.expect_tx (+ tx bytes of both CTAs' loads)
            mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 2 * Cfg::EPI_WARPS);
        }
        for (int w = 0; w < Cfg::EPI_WARPS; ++w) mbar_init(&res_bar[w], 1);
        if (TMARES) tma_prefetch_desc(&tmR);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_cg2<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tcgen05_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();  // everything above overlapped the previous kernel's tail; its outputs are visible from here on

    const int num_super = p.m_pairs * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows + own half of the W tile, completion on the LEADER's barrier =====
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int st = cluster_id; st < num_super; st += num_clusters) {
                const int m_blk = (st / p.n_tiles) * 2 + (int)cta_rank, n_blk = st % p.n_tiles;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait_cluster(&empty_bar[stage], phase ^ 1);
                    if ((p.debug & 1) && (phase || st != cluster_id)) {
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 0);
                    } else {
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        const bool same = (p.debug & 2) != 0;  // experiment: always the same (L2-resident) boxes
                        tma_load_2d_cg2(sA + stage * A_STAGE_BYTES, &tmA, fb, same ? 0 : kb * BK,
                                        same ? (int)cta_rank * BM : m_blk * BM);
                        tma_load_2d_cg2(sB + stage * Cfg::B_HALF_BYTES, &tmB, fb, same ? 0 : kb * BK,
                                        (same ? 0 : n_blk * BN) + (int)cta_rank * (BN / 2));
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (cta_rank == 0 && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int st = cluster_id; st < num_super; st += num_clusters) {
                mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait_cluster(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_HALF_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_bf16_ss_cg2(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                         (uint32_t)((kb | k) != 0));
                    }
                    umma_commit_mc_cg2(&empty_bar[stage], (uint16_t)0x3);  // frees the stage in both CTAs
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_mc_cg2(&tmem_full_bar[as], (uint16_t)0x3);  // accumulator ready in both CTAs
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else {
        // ===== epilogue warps (both CTAs, each on its own 128 rows) =====
        const int ew = warp - 2;
        const int q = warp & 3;     // TMEM lane quadrant this warp may read
        const int cgp = ew >> 2;    // which 64-column group of the tile
        const bool has_res = ANY ? p.residual != nullptr : (EF & EF_RES) != 0;
        const bool ln = ANY ? p.ln_stats != nullptr : (EF & EF_LN) != 0;
        float* myBias = sBias + ew * Cfg::COLS_PER_WARP;
        uint8_t* myStage = sOutStage + ew * Cfg::OUT_TILE_BYTES;
        bool store_pending = false;
        uint32_t res_phase = 0;
        int as = 0;
        uint32_t aphase = 0;
        for (int st = cluster_id; st < num_super; st += num_clusters) {
            const int m_blk = (st / p.n_tiles) * 2 + (int)cta_rank, n_blk = st % p.n_tiles;
            const int n0 = n_blk * BN + cgp * Cfg::COLS_PER_WARP;
            const int row0 = m_blk * BM + q * 32;
            for (int c = lane; c < Cfg::COLS_PER_WARP; c += 32) {
                myBias[c] = (p.bias != nullptr && n0 + c < p.N) ? p.bias[n0 + c] : 0.0f;
                if (ln) myBias[Cfg::EPI_WARPS * Cfg::COLS_PER_WARP + c] = (n0 + c < p.N) ? p.ln_c1[n0 + c] : 0.0f;
            }
            __syncwarp();
            uint4 rv[4];
            float ln_rstd = 1.0f, ln_mr = 0.0f;
            if (has_res && !TMARES) load_residual_chunk(p, row0 + lane, n0, rv);
            if (ln) load_ln_row(p, row0 + lane, ln_rstd, ln_mr);
            mbar_wait_cluster(&tmem_full_bar[as], aphase);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + cgp * Cfg::COLS_PER_WARP);
            epilogue_tile<BN, EF>(p, myBias, t_addr, row0, n0, lane, rv, ln_rstd, ln_mr, &tmC, myStage, store_pending, &tmR,
                                  &res_bar[ew], res_phase);
            tcgen05_fence_before();
            __syncwarp();
            // the accumulator values are in registers (tcgen05.wait::ld): a relaxed arrive is enough, and unlike a
            // release at cluster scope it does not wait for this warp's output stores to drain
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (lane == 0) tma_store_wait_all();  // staged tiles fully written out before the CTA's smem goes away
    }

    // no CTA may exit while its peer can still read its smem or arrive on its barriers
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_cg2<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int BN, uint32_t EF>
static int launch_gemm_ef(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                          const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "GEMM shared memory budget");
    GVL_CUDA(cudaFuncSetAttribute(gemm_bf16_cg2_kernel<BN, EF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    const int super_tiles = p.m_pairs * p.n_tiles;
    const int max_clusters = sm_count() / 2;
    const int clusters = super_tiles < max_clusters ? super_tiles : max_clusters;
    ProfScope prof(GVL_K_GEMM, 2.0 * p.M * (double)p.N * p.K, stream);
    GVL_CUDA(launch_pdl(gemm_bf16_cg2_kernel<BN, EF>, dim3(2 * clusters), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, stream, tmA,
                        tmB, tmC, tmR, p));
    GVL_LAUNCH_CHECK("gemm_bf16_cg2_kernel");
    return 0;
}

// the flag combinations of the tower's hot GEMMs get their own instantiation; everything else is generic
template <int BN>
static int launch_gemm_bn(uint32_t ef, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const CUtensorMap& tmR, const GemmParams& p, cudaStream_t stream) {
    switch (ef) {
        case 0u: return launch_gemm_ef<BN, 0u>(tmA, tmB, tmC, tmR, p, stream);                                  // qkv, kv
        case GVL_ACT_GELU_TANH: return launch_gemm_ef<BN, GVL_ACT_GELU_TANH>(tmA, tmB, tmC, tmR, p, stream);    // fc1
        case GVL_ACT_GELU_ERF: return launch_gemm_ef<BN, GVL_ACT_GELU_ERF>(tmA, tmB, tmC, tmR, p, stream);      // VideoMAE fc1
        case EF_RES: return launch_gemm_ef<BN, EF_RES>(tmA, tmB, tmC, tmR, p, stream);                          // out, fc2, patch
        case EF_RES | EF_STATS: return launch_gemm_ef<BN, EF_RES | EF_STATS>(tmA, tmB, tmC, tmR, p, stream);    // fold_ln producers
        case EF_RES | EF_TMARES: return launch_gemm_ef<BN, EF_RES | EF_TMARES>(tmA, tmB, tmC, tmR, p, stream);
        case EF_RES | EF_STATS | EF_TMARES:
            return launch_gemm_ef<BN, EF_RES | EF_STATS | EF_TMARES>(tmA, tmB, tmC, tmR, p, stream);
        case EF_LN: return launch_gemm_ef<BN, EF_LN>(tmA, tmB, tmC, tmR, p, stream);                            // fold_ln qkv, kv
        case EF_LN | GVL_ACT_GELU_TANH:
            return launch_gemm_ef<BN, EF_LN | GVL_ACT_GELU_TANH>(tmA, tmB, tmC, tmR, p, stream);                // fold_ln fc1
        default: return launch_gemm_ef<BN, EF_ANY>(tmA, tmB, tmC, tmR, p, stream);
    }
}

static int pick_bn(int N) {
    // smallest padded width wins; ties go to the wider tile (fewer A re-reads, more MMA per barrier).
    const int cands[3] = {256, 192, 128};
    int best = 256;
    long best_pad = -1;
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const long padded = (long)((N + bn - 1) / bn) * bn;
        if (best_pad < 0 || padded < best_pad) {
            best_pad = padded;
            best = bn;
        }
    }
    return best;
}

}  // namespace gvl

extern "C" int gvl_gemm_stats_slots(int N) {
    // one slot per epilogue warp column group (64 columns), padded to an even count
    const int bn = gvl::pick_bn(N);
    const int slots = (bn / 64) * ((N + bn - 1) / bn);
    return (slots + 1) & ~1;
}

extern "C" int gvl_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                             int ldr, int res_row_mod, void* out, int ldo, int out_f32, int M, int N, int K, int act,
                             void* stream) {
    return gvl_gemm_bf16_fused(A, lda, W, ldw, bias, residual, ldr, res_row_mod, out, ldo, out_f32, M, N, K, act, nullptr,
                               stream);
}

extern "C" int gvl_gemm_bf16_fused(const void* A, int lda, const void* W, int ldw, const float* bias,
                                   const void* residual, int ldr, int res_row_mod, void* out, int ldo, int out_f32, int M,
                                   int N, int K, int act, const gvl_gemm_fusion* fusion, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(A && W && out, "gvl_gemm_bf16: null pointer");
    GVL_CHECK_ARG(M > 0 && N > 0 && K > 0, "gvl_gemm_bf16: bad shape M=%d N=%d K=%d", M, N, K);
    GVL_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "gvl_gemm_bf16: N and K must be multiples of 8 (N=%d K=%d)", N, K);
    GVL_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "gvl_gemm_bf16: bad lda/ldw %d/%d", lda, ldw);
    GVL_CHECK_ARG(ldo >= N && ldo % (out_f32 ? 4 : 8) == 0, "gvl_gemm_bf16: bad ldo %d", ldo);
    GVL_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)out % 16 == 0),
                  "gvl_gemm_bf16: pointers must be 16-byte aligned");
    GVL_CHECK_ARG(residual == nullptr || (ldr % 8 == 0 && ldr >= N && (uintptr_t)residual % 16 == 0),
                  "gvl_gemm_bf16: bad residual ld/alignment");
    GVL_CHECK_ARG(act >= 0 && act <= 2, "gvl_gemm_bf16: bad act %d", act);

    int bn = pick_bn(N);
    static const int debug = [] { const char* e = getenv("GVL_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
    static const int force_bn = [] { const char* e = getenv("GVL_GEMM_BN"); return e ? atoi(e) : 0; }();
    if ((force_bn == 256 || force_bn == 192 || force_bn == 128) && fusion == nullptr) bn = force_bn;
    GemmParams p;
    p.debug = debug;
    p.stats_out = nullptr;
    p.stats_slots = gvl_gemm_stats_slots(N);
    p.ln_stats = nullptr;
    p.ln_c1 = nullptr;
    p.ln_slots = 0;
    p.ln_inv_d = p.ln_eps = 0.f;
    if (fusion != nullptr) {
        GVL_CHECK_ARG(fusion->stats_out == nullptr || !out_f32, "gvl_gemm_bf16_fused: row statistics need a bf16 output");
        GVL_CHECK_ARG(fusion->ln_stats == nullptr || (fusion->ln_c1 && fusion->ln_slots >= 0 && fusion->ln_dim > 0),
                      "gvl_gemm_bf16_fused: incomplete LayerNorm fusion arguments");
        GVL_CHECK_ARG(fusion->ln_stats == nullptr ||
                          (fusion->ln_slots % 2 == 0 && fusion->ln_slots <= 2 * kMaxLnSlotPairs &&
                           (uintptr_t)fusion->ln_stats % (fusion->ln_slots ? 16 : 8) == 0),
                      "gvl_gemm_bf16_fused: ln_slots must be even and <= %d, ln_stats 16-byte aligned", 2 * kMaxLnSlotPairs);
        p.stats_out = fusion->stats_out;
        p.ln_stats = fusion->ln_stats;
        p.ln_c1 = fusion->ln_c1;
        p.ln_slots = fusion->ln_slots;
        p.ln_inv_d = fusion->ln_dim > 0 ? 1.0f / (float)fusion->ln_dim : 0.f;
        p.ln_eps = fusion->ln_eps;
    }
    p.bias = bias;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.ldr = ldr;
    p.res_row_mod = res_row_mod;
    p.out = out;
    p.ldo = ldo;
    p.out_f32 = out_f32;
    p.M = M;
    p.N = N;
    p.K = K;
    p.act = act;
    p.m_pairs = ((M + BM - 1) / BM + 1) / 2;
    p.n_tiles = (N + bn - 1) / bn;
    p.k_blocks = (K + BK - 1) / BK;

    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK, true);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(bn / 2), BK, true);
    if (rc) return rc;

    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    // output map for the epilogue's TMA stores: 32 x 32 bf16 boxes, SWIZZLE_64B staging tiles
    CUtensorMap tmC = tmA;  // placeholder for fp32 outputs (never dereferenced)
    if (!out_f32) {
        const uint64_t cdims[2] = {(uint64_t)N, (uint64_t)M};
        const uint64_t cstr[1] = {(uint64_t)ldo * 2};
        const uint32_t cbox[2] = {32, 32};
        rc = make_tmap_nd_bf16(&tmC, out, 2, cdims, cstr, cbox, 64);
        if (rc) return rc;
    }
    uint32_t ef = (uint32_t)act | (residual ? EF_RES : 0u) | (p.ln_stats ? EF_LN : 0u) |
                  (p.stats_out ? EF_STATS : 0u) | (out_f32 ? EF_F32 : 0u);
    // residual tile through TMA (GVL_GEMM_TMARES=0 turns it off for A/B runs): plain row-for-row bf16 residuals of the
    // two hot flavours only; out-proj 111 -> 103 us, fc2 350 -> 343 us
    static const bool tma_res = [] { const char* e = getenv("GVL_GEMM_TMARES"); return !(e && e[0] == '0'); }();
    CUtensorMap tmR = tmC;
    if (tma_res && residual && res_row_mod == 0 && !out_f32 && (ef == EF_RES || ef == (EF_RES | EF_STATS))) {
        const uint64_t rdims[2] = {(uint64_t)N, (uint64_t)M};
        const uint64_t rstr[1] = {(uint64_t)ldr * 2};
        const uint32_t rbox[2] = {32, 32};
        rc = make_tmap_nd_bf16(&tmR, residual, 2, rdims, rstr, rbox, 64);
        if (rc) return rc;
        ef |= EF_TMARES;
    }
    switch (bn) {
        case 256: return launch_gemm_bn<256>(ef, tmA, tmB, tmC, tmR, p, s);
        case 192: return launch_gemm_bn<192>(ef, tmA, tmB, tmC, tmR, p, s);
        default: return launch_gemm_bn<128>(ef, tmA, tmB, tmC, tmR, p, s);
    }
}
