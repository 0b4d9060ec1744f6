// siglip.cu — K6/K7: the whole SigLIP2 vision tower as one stream-ordered call, and the projector.
//
// The composite only sequences the kernels of this library on the caller's stream (no allocation,
// no synchronisation): patch GEMM (+bias +position embedding) -> L x [LN, QKV GEMM, attention,
// out-proj GEMM (+residual), LN, fc1 GEMM (+GELU), fc2 GEMM (+residual)] -> post-LN -> MAP head.
// The residual stream is updated in place by the GEMM epilogues.
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace gvl {

struct Workspace {
    uint8_t* base;
    size_t off, cap;
    void* take(size_t bytes) {
        size_t a = (off + 255) & ~(size_t)255;
        off = a + bytes;
        return base ? base + a : nullptr;
    }
};

struct VitBuffers {
    void *x, *xn, *qkv, *attn, *h, *pa, *hx, *hxn, *hh;
    void* tiles;  // query-tile table of a ragged batch (gvl_attention_varlen_tiles), 16 B per tile
    float *stats_a, *stats_b;  // LayerNorm-fusion partial row sums ([M, slots, 2] each)
    float *fin_a, *fin_b;      // finalised (rstd, mean * rstd) per row ([M, 2] each)
};

static size_t carve(const gvl_vit_weights* w, size_t M, int B, uint8_t* base, VitBuffers& vb) {
    const size_t D = w->D, I = w->I;
    Workspace ws{base, 0, 0};
    vb.x = ws.take(M * D * 2);
    vb.xn = ws.take(M * D * 2);
    vb.qkv = ws.take(M * 3 * D * 2);
    vb.attn = ws.take(M * D * 2);
    vb.h = ws.take(M * I * 2);
    vb.pa = ws.take((size_t)B * D * 2);
    vb.hx = ws.take((size_t)B * D * 2);
    vb.hxn = ws.take((size_t)B * D * 2);
    vb.hh = ws.take((size_t)B * I * 2);
    const size_t slots = (size_t)gvl_gemm_stats_slots((int)D);
    vb.stats_a = reinterpret_cast<float*>(ws.take(M * slots * 2 * sizeof(float)));
    vb.stats_b = reinterpret_cast<float*>(ws.take(M * slots * 2 * sizeof(float)));
    vb.fin_a = reinterpret_cast<float*>(ws.take(M * 2 * sizeof(float)));
    vb.fin_b = reinterpret_cast<float*>(ws.take(M * 2 * sizeof(float)));
    vb.tiles = ws.take((M / 128 + (size_t)B + 1) * 16);  // sum ceil(T_i / 128) <= M / 128 + B
    return ws.off + 256;
}

}  // namespace gvl

extern "C" size_t gvl_siglip_workspace_bytes(const gvl_vit_weights* w, int B) {
    if (!w || B <= 0) return 0;
    gvl::VitBuffers vb;
    return gvl::carve(w, (size_t)B * w->T, B, nullptr, vb);
}

extern "C" size_t gvl_siglip_ragged_workspace_bytes(const gvl_vit_weights* w, long long M_total, int B_total) {
    if (!w || M_total <= 0 || B_total <= 0) return 0;
    gvl::VitBuffers vb;
    return gvl::carve(w, (size_t)M_total, B_total, nullptr, vb);
}

#define GVL_TRY(call)          \
    do {                       \
        int rc__ = (call);     \
        if (rc__) return rc__; \
    } while (0)

// The tower over a ragged batch: `n_groups` groups, group g = B_g items of T_g tokens each with its own position table,
// token rows concatenated in group order (M rows, B items in total).  Everything row-wise (GEMMs, LayerNorms,
// statistics) runs ONCE over all M rows — the weights do not care where an item ends — and only the three item-aware
// steps run per group: the patch GEMM (position add, row % T_g), the attention and the MAP-head probe attention.
static int forward_groups(const gvl_vit_weights* w, const void* patches, int n_groups, const gvl_ragged_group* groups,
                          void* workspace, size_t workspace_bytes, void* pooled, void* last_hidden, void* stream,
                          const char* who) {
    using namespace gvl;
    GVL_CHECK_ARG(w->D == w->H * w->hd && w->L > 0 && w->layers, "%s: inconsistent weight pack", who);
    GVL_CHECK_ARG((uintptr_t)workspace % 256 == 0, "%s: workspace must be 256-byte aligned", who);
    long long M_ll = 0, B_ll = 0;
    for (int g = 0; g < n_groups; ++g) {
        GVL_CHECK_ARG(groups[g].B > 0 && groups[g].T > 0 && groups[g].pos, "%s: bad group %d (B=%d T=%d)", who, g,
                      groups[g].B, groups[g].T);
        M_ll += (long long)groups[g].B * groups[g].T;
        B_ll += groups[g].B;
    }
    GVL_CHECK_ARG(M_ll <= 2147483647LL / w->I, "%s: %lld token rows exceed the 32-bit index range", who, M_ll);
    const int M = (int)M_ll, B = (int)B_ll;
    VitBuffers vb;
    const size_t need = carve(w, (size_t)M, B, reinterpret_cast<uint8_t*>(workspace), vb);
    GVL_CHECK_ARG(workspace_bytes >= need, "%s: workspace %zu < required %zu bytes", who, workspace_bytes, need);

    const int D = w->D, I = w->I, H = w->H, hd = w->hd;
    const float scale = 1.0f / sqrtf((float)hd);
    auto rows = [](void* p, size_t row, size_t ld_elems) { return reinterpret_cast<uint8_t*>(p) + row * ld_elems * 2; };
    auto crows = [](const void* p, size_t row, size_t ld_elems) {
        return reinterpret_cast<const uint8_t*>(p) + row * ld_elems * 2;
    };
    // per-group steps
    auto patch_embed = [&](const gvl_gemm_fusion* prod) -> int {
        size_t r0 = 0;
        for (int g = 0; g < n_groups; ++g) {
            const int Mg = groups[g].B * groups[g].T;
            gvl_gemm_fusion pg;
            if (prod) {
                pg = *prod;
                pg.stats_out = prod->stats_out + r0 * (size_t)gvl_gemm_stats_slots(D) * 2;
            }
            GVL_TRY(gvl_gemm_bf16_fused(crows(patches, r0, w->patch_ld), w->patch_ld, w->w_patch, w->patch_ld, w->b_patch,
                                        groups[g].pos, D, groups[g].T, rows(vb.x, r0, D), D, 0, Mg, D, w->patch_ld,
                                        GVL_ACT_NONE, prod ? &pg : nullptr, stream));
            r0 += Mg;
        }
        return 0;
    };
    // a ragged batch takes ONE attention launch per layer over a table of query tiles (uploaded once per call)
    int n_tiles = 0;
    double score_elems = 0.0;
    if (n_groups > 1) {
        std::vector<int32_t> item_tokens;
        for (int g = 0; g < n_groups; ++g) {
            item_tokens.insert(item_tokens.end(), (size_t)groups[g].B, groups[g].T);
            score_elems += (double)groups[g].B * groups[g].T * (double)groups[g].T;
        }
        GVL_TRY(gvl_attention_varlen_tiles((int)item_tokens.size(), item_tokens.data(), nullptr, &n_tiles));
        std::vector<int32_t> table((size_t)n_tiles * 4);
        GVL_TRY(gvl_attention_varlen_tiles((int)item_tokens.size(), item_tokens.data(), table.data(), &n_tiles));
        // pageable source: staged before the call returns
        GVL_CUDA(cudaMemcpyAsync(vb.tiles, table.data(), table.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 reinterpret_cast<cudaStream_t>(stream)));
    }
    auto attention = [&]() -> int {
        if (n_groups > 1)
            return gvl_attention_varlen_bf16(vb.qkv, vb.attn, M, vb.tiles, n_tiles, score_elems, H, hd, scale, stream);
        size_t r0 = 0;
        for (int g = 0; g < n_groups; ++g) {
            GVL_TRY(gvl_attention_bf16(rows(vb.qkv, r0, 3 * (size_t)D), rows(vb.attn, r0, D), groups[g].B, groups[g].T, H, hd,
                                       scale, stream));
            r0 += (size_t)groups[g].B * groups[g].T;
        }
        return 0;
    };
    auto probe = [&]() -> int {
        size_t r0 = 0, b0 = 0;
        for (int g = 0; g < n_groups; ++g) {
            GVL_TRY(gvl_probe_attention_bf16(w->probe_q, rows(vb.qkv, r0, 2 * (size_t)D), rows(vb.pa, b0, D), groups[g].B,
                                             groups[g].T, H, hd, stream));
            r0 += (size_t)groups[g].B * groups[g].T;
            b0 += groups[g].B;
        }
        return 0;
    };

    if (w->fold_ln) {
        // LayerNorm folded into the GEMMs (gvl_gemm_fusion): the GEMMs that write the residual stream x also write
        // its partial row sums; the GEMMs that consume LayerNorm(x) read x and normalise in their epilogue.
        GVL_CHECK_ARG(w->c1_kv != nullptr, "%s: fold_ln pack without c1 vectors", who);
        const int slots = gvl_gemm_stats_slots(D);
        // the partial sums are reduced to (rstd, mean * rstd) once per row by a small kernel and the consumers read
        // 8 bytes per row instead of the slabs (identical bits; +2.6 % in step; GVL_LN_FINALIZE=0 for A/B runs)
        static const bool finalize = [] { const char* e = getenv("GVL_LN_FINALIZE"); return !(e && e[0] == '0'); }();
        gvl_gemm_fusion prod_a = {vb.stats_a, nullptr, 0, 0, nullptr, 0.f};
        gvl_gemm_fusion prod_b = {vb.stats_b, nullptr, 0, 0, nullptr, 0.f};
        auto consumer = [&](float* stats, float* fin, const float* c1, gvl_gemm_fusion& f) -> int {
            if (finalize) {
                int rc = gvl_ln_finalize(stats, M, slots, D, w->eps, fin, stream);
                if (rc) return rc;
                f = gvl_gemm_fusion{nullptr, fin, 0, D, c1, w->eps};
            } else {
                f = gvl_gemm_fusion{nullptr, stats, slots, D, c1, w->eps};
            }
            return 0;
        };
        GVL_TRY(patch_embed(&prod_a));
        for (int l = 0; l < w->L; ++l) {
            const gvl_vit_layer& ly = w->layers[l];
            gvl_gemm_fusion ln1, ln2;
            GVL_TRY(consumer(vb.stats_a, vb.fin_a, ly.c1_qkv, ln1));
            GVL_TRY(gvl_gemm_bf16_fused(vb.x, D, ly.w_qkv, D, ly.b_qkv, nullptr, 0, 0, vb.qkv, 3 * D, 0, M, 3 * D, D,
                                        GVL_ACT_NONE, &ln1, stream));
            GVL_TRY(attention());
            GVL_TRY(gvl_gemm_bf16_fused(vb.attn, D, ly.w_o, D, ly.b_o, vb.x, D, 0, vb.x, D, 0, M, D, D, GVL_ACT_NONE,
                                        &prod_b, stream));
            GVL_TRY(consumer(vb.stats_b, vb.fin_b, ly.c1_fc1, ln2));
            GVL_TRY(gvl_gemm_bf16_fused(vb.x, D, ly.w_fc1, D, ly.b_fc1, nullptr, 0, 0, vb.h, I, 0, M, I, D, w->act, &ln2,
                                        stream));
            GVL_TRY(gvl_gemm_bf16_fused(vb.h, I, ly.w_fc2, I, ly.b_fc2, vb.x, D, 0, vb.x, D, 0, M, D, I, GVL_ACT_NONE,
                                        &prod_a, stream));
        }
        if (last_hidden)  // only materialised when the caller asks for the post-layernorm tokens
            GVL_TRY(gvl_layernorm_bf16(vb.x, D, w->post_g, w->post_b, last_hidden, D, M, D, w->eps, stream));
        gvl_gemm_fusion lnp;
        GVL_TRY(consumer(vb.stats_a, vb.fin_a, w->c1_kv, lnp));
        GVL_TRY(gvl_gemm_bf16_fused(vb.x, D, w->w_kv, D, w->b_kv, nullptr, 0, 0, vb.qkv, 2 * D, 0, M, 2 * D, D,
                                    GVL_ACT_NONE, &lnp, stream));
    } else {
        // embeddings: conv-as-GEMM + bias + learned position embedding (row % T)
        GVL_TRY(patch_embed(nullptr));
        for (int l = 0; l < w->L; ++l) {
            const gvl_vit_layer& ly = w->layers[l];
            GVL_TRY(gvl_layernorm_bf16(vb.x, D, ly.ln1_g, ly.ln1_b, vb.xn, D, M, D, w->eps, stream));
            GVL_TRY(gvl_gemm_bf16(vb.xn, D, ly.w_qkv, D, ly.b_qkv, nullptr, 0, 0, vb.qkv, 3 * D, 0, M, 3 * D, D,
                                  GVL_ACT_NONE, stream));
            GVL_TRY(attention());
            GVL_TRY(gvl_gemm_bf16(vb.attn, D, ly.w_o, D, ly.b_o, vb.x, D, 0, vb.x, D, 0, M, D, D, GVL_ACT_NONE, stream));
            GVL_TRY(gvl_layernorm_bf16(vb.x, D, ly.ln2_g, ly.ln2_b, vb.xn, D, M, D, w->eps, stream));
            GVL_TRY(gvl_gemm_bf16(vb.xn, D, ly.w_fc1, D, ly.b_fc1, nullptr, 0, 0, vb.h, I, 0, M, I, D, w->act, stream));
            GVL_TRY(gvl_gemm_bf16(vb.h, I, ly.w_fc2, I, ly.b_fc2, vb.x, D, 0, vb.x, D, 0, M, D, I, GVL_ACT_NONE, stream));
        }
        void* tokens = last_hidden ? last_hidden : vb.xn;
        GVL_TRY(gvl_layernorm_bf16(vb.x, D, w->post_g, w->post_b, tokens, D, M, D, w->eps, stream));

        // MAP head: K/V projections of all tokens, probe attention, out-proj, LN, MLP with residual
        GVL_TRY(gvl_gemm_bf16(tokens, D, w->w_kv, D, w->b_kv, nullptr, 0, 0, vb.qkv, 2 * D, 0, M, 2 * D, D, GVL_ACT_NONE,
                              stream));
    }
    GVL_TRY(probe());
    GVL_TRY(gvl_gemm_bf16(vb.pa, D, w->w_ho, D, w->b_ho, nullptr, 0, 0, vb.hx, D, 0, B, D, D, GVL_ACT_NONE, stream));
    GVL_TRY(gvl_layernorm_bf16(vb.hx, D, w->hln_g, w->hln_b, vb.hxn, D, B, D, w->eps, stream));
    GVL_TRY(gvl_gemm_bf16(vb.hxn, D, w->w_hfc1, D, w->b_hfc1, nullptr, 0, 0, vb.hh, I, 0, B, I, D, w->act, stream));
    GVL_TRY(gvl_gemm_bf16(vb.hh, I, w->w_hfc2, I, w->b_hfc2, vb.hx, D, 0, pooled, D, 0, B, D, I, GVL_ACT_NONE, stream));
    return 0;
}

extern "C" int gvl_siglip_forward(const gvl_vit_weights* w, const void* patches, int B, void* workspace,
                                  size_t workspace_bytes, void* pooled, void* last_hidden, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(w && patches && workspace && pooled, "gvl_siglip_forward: null pointer");
    GVL_CHECK_ARG(B > 0, "gvl_siglip_forward: bad batch %d", B);
    const gvl_ragged_group one = {B, w->T, w->pos};
    return forward_groups(w, patches, 1, &one, workspace, workspace_bytes, pooled, last_hidden, stream, "gvl_siglip_forward");
}

extern "C" int gvl_siglip_forward_ragged(const gvl_vit_weights* w, const void* patches, int n_groups,
                                         const gvl_ragged_group* groups, void* workspace, size_t workspace_bytes,
                                         void* pooled, void* last_hidden, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(w && patches && workspace && pooled && groups, "gvl_siglip_forward_ragged: null pointer");
    GVL_CHECK_ARG(n_groups > 0 && n_groups <= 4096, "gvl_siglip_forward_ragged: bad group count %d", n_groups);
    return forward_groups(w, patches, n_groups, groups, workspace, workspace_bytes, pooled, last_hidden, stream,
                          "gvl_siglip_forward_ragged");
}

extern "C" int gvl_project(const void* x, int M, int enc_dim, int llm_dim, const void* w1, const float* b1,
                           const void* w2, const float* b2, void* hidden, void* out, int out_f32, void* stream) {
    using namespace gvl;
    GVL_CHECK_ARG(x && w1 && w2 && hidden && out, "gvl_project: null pointer");
    GVL_TRY(gvl_gemm_bf16(x, enc_dim, w1, enc_dim, b1, nullptr, 0, 0, hidden, llm_dim, 0, M, llm_dim, enc_dim,
                          GVL_ACT_GELU_ERF, stream));
    GVL_TRY(gvl_gemm_bf16(hidden, llm_dim, w2, llm_dim, b2, nullptr, 0, 0, out, llm_dim, out_f32, M, llm_dim, llm_dim,
                          GVL_ACT_NONE, stream));
    return 0;
}
