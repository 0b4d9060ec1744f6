// videomae.cu — the VideoMAE clip encoder as one stream-ordered call through the same
// kernels as the SigLIP tower: tubelet patch GEMM (+bias +fixed sinusoid position table) -> L x [LN, QKV GEMM,
// attention, out-proj GEMM (+residual), LN, fc1 GEMM (+GELU erf), fc2 GEMM (+residual)] -> optional final LN ->
// mean over the tokens (the reference's `outputs.last_hidden_state.mean(dim=1)`, scripts/extract_features.py:381).
#include "common.cuh"

namespace gvl {

// x: bf16 [B, T, D] -> out [B, D] = mean over T (fp32 accumulation).  grid (D / 256 column slabs, B); the 8 warps of a
// CTA take interleaved token rows of the slab (lane = 8 consecutive columns, 16-byte loads) and reduce through smem.
__global__ void __launch_bounds__(256)
mean_tokens_kernel(const __nv_bfloat16* __restrict__ x, int T, int D, void* __restrict__ out, int out_f32) {
    __shared__ float part[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const int b = blockIdx.y;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 < D) {
        const __nv_bfloat16* base = x + (size_t)b * T * D + c0;
        for (int t = warp; t < T; t += 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)t * D);
            acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x);
            acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
            acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z);
            acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
        s /= (float)T;
        if (out_f32)
            reinterpret_cast<float*>(out)[(size_t)b * D + c] = s;
        else
            reinterpret_cast<__nv_bfloat16*>(out)[(size_t)b * D + c] = __float2bfloat16_rn(s);
    }
}

struct VmaeBuffers {
    void *x, *xn, *qkv, *attn, *h;
};

static size_t vmae_carve(const gvl_vit_weights* w, int B, uint8_t* base, VmaeBuffers& vb) {
    const size_t M = (size_t)B * w->T, D = w->D, I = w->I;
    size_t off = 0;
    auto take = [&](size_t bytes) -> void* {
        const size_t a = (off + 255) & ~(size_t)255;
        off = a + bytes;
        return base ? base + a : nullptr;
    };
    vb.x = take(M * D * 2);
    vb.xn = take(M * D * 2);
    vb.qkv = take(M * 3 * D * 2);
    vb.attn = take(M * D * 2);
    vb.h = take(M * I * 2);
    return off + 256;
}

}  // namespace gvl

#define GVL_TRY(call)          \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

extern "C" int gvl_mean_tokens_bf16(const void* x, int B, int T, int D, void* out, int out_f32, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(x && out, "gvl_mean_tokens_bf16: null pointer");
    GVL_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && D > 0 && D % 8 == 0, "gvl_mean_tokens_bf16: bad shape B=%d T=%d D=%d", B,
                  T, D);
    GVL_CHECK_ARG((uintptr_t)x % 16 == 0, "gvl_mean_tokens_bf16: x must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    ProfScope prof(GVL_K_LAYERNORM, (double)B * T * D * 2, s);
    mean_tokens_kernel<<<dim3((D + 255) / 256, B), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), T, D, out,
                                                              out_f32);
    GVL_LAUNCH_CHECK("mean_tokens_kernel");
    return 0;
}

extern "C" size_t gvl_videomae_workspace_bytes(const gvl_vit_weights* w, int B) {
    if (!w || B <= 0) return 0;
    gvl::VmaeBuffers vb;
    return gvl::vmae_carve(w, B, nullptr, vb);
}

extern "C" int gvl_videomae_forward(const gvl_vit_weights* w, const void* patches, int B, void* workspace,
                                    size_t workspace_bytes, void* pooled, int pooled_f32, void* last_hidden,
                                    void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(w && patches && workspace && pooled, "gvl_videomae_forward: null pointer");
    GVL_CHECK_ARG(B > 0, "gvl_videomae_forward: bad batch %d", B);
    GVL_CHECK_ARG(w->D == w->H * w->hd && w->L > 0 && w->layers, "gvl_videomae_forward: inconsistent weight pack");
    GVL_CHECK_ARG((uintptr_t)workspace % 256 == 0, "gvl_videomae_forward: workspace must be 256-byte aligned");
    VmaeBuffers vb;
    const size_t need = vmae_carve(w, B, reinterpret_cast<uint8_t*>(workspace), vb);
    GVL_CHECK_ARG(workspace_bytes >= need, "gvl_videomae_forward: workspace %zu < required %zu bytes", workspace_bytes,
                  need);
    const int D = w->D, I = w->I, T = w->T, H = w->H, hd = w->hd;
    const int M = B * T;
    const float scale = 1.0f / sqrtf((float)hd);

    // tubelet embedding: Conv3d-as-GEMM + bias + fixed sinusoid position table (row % T)
    GVL_TRY(gvl_gemm_bf16(patches, w->patch_ld, w->w_patch, w->patch_ld, w->b_patch, w->pos, D, T, vb.x, D, 0, M, D,
                          w->patch_ld, GVL_ACT_NONE, stream));
    for (int l = 0; l < w->L; ++l) {
        const gvl_vit_layer& ly = w->layers[l];
        GVL_TRY(gvl_layernorm_bf16(vb.x, D, ly.ln1_g, ly.ln1_b, vb.xn, D, M, D, w->eps, stream));
        GVL_TRY(gvl_gemm_bf16(vb.xn, D, ly.w_qkv, D, ly.b_qkv, nullptr, 0, 0, vb.qkv, 3 * D, 0, M, 3 * D, D,
                              GVL_ACT_NONE, stream));
        GVL_TRY(gvl_attention_bf16(vb.qkv, vb.attn, B, T, H, hd, scale, stream));
        GVL_TRY(gvl_gemm_bf16(vb.attn, D, ly.w_o, D, ly.b_o, vb.x, D, 0, vb.x, D, 0, M, D, D, GVL_ACT_NONE, stream));
        GVL_TRY(gvl_layernorm_bf16(vb.x, D, ly.ln2_g, ly.ln2_b, vb.xn, D, M, D, w->eps, stream));
        GVL_TRY(gvl_gemm_bf16(vb.xn, D, ly.w_fc1, D, ly.b_fc1, nullptr, 0, 0, vb.h, I, 0, M, I, D, w->act, stream));
        GVL_TRY(gvl_gemm_bf16(vb.h, I, ly.w_fc2, I, ly.b_fc2, vb.x, D, 0, vb.x, D, 0, M, D, I, GVL_ACT_NONE, stream));
    }
    const void* tokens = vb.x;
    if (w->post_g != nullptr) {  // checkpoints with use_mean_pooling = false carry a final LayerNorm
        void* dst = last_hidden ? last_hidden : vb.xn;
        GVL_TRY(gvl_layernorm_bf16(vb.x, D, w->post_g, w->post_b, dst, D, M, D, w->eps, stream));
        tokens = dst;
    } else if (last_hidden) {
        GVL_CUDA(cudaMemcpyAsync(last_hidden, vb.x, (size_t)M * D * 2, cudaMemcpyDeviceToDevice,
                                 reinterpret_cast<cudaStream_t>(stream)));
    }
    GVL_TRY(gvl_mean_tokens_bf16(tokens, B, T, D, pooled, pooled_f32, stream));
    return 0;
}
