// Host side of libmmf_b200.so: checkpoint packing, packed-layout planning, per-step kernel schedule and
// the extern "C" entry points declared in include/mmf_b200.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mmf_b200.h"
#include "mmf_epic.h"
#include "mmf_host.h"
#include "mmf_internal.h"
#include "mmf_simt.h"
#include "mmf_tftile.h"

namespace mmf {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

// ---------------------------------------------------------------------------------------------
// packed parameters
// ---------------------------------------------------------------------------------------------
struct BlockOff {               // arena offsets of one (possibly 2-group) attention block
    int groups = 1, C = 256;
    size_t wqkv, wproj, wfc, wp2, bqkv, bproj, bfc, bp2, ln1g, ln1b, ln2g, ln2b, qg, qb, kg, kb;
    bool qkln = true;
};
struct BlockDev {
    int groups, C, hs;
    const bf16 *wqkv, *wproj, *wfc, *wp2;
    const float *bqkv, *bproj, *bfc, *bp2, *ln1g, *ln1b, *ln2g, *ln2b, *qg, *qb, *kg, *kb;
    CUtensorMap tm_wqkv, tm_wproj, tm_wfc, tm_wp2;
};

struct Workspace {
    int mcap = 0;                // packed-row capacity, multiple of 128
    int slot_cap = 0;            // B*D capacity of the staging buffers
    int tcap = 0;                // rows of the time tables
    int item_cap = 0;
    uint8_t* base = nullptr;
    float *resid = nullptr, *skip = nullptr, *xs = nullptr, *temb = nullptr, *temb2 = nullptr;
    bf16 *act = nullptr, *q = nullptr, *k = nullptr, *vt = nullptr, *attn = nullptr, *hidden = nullptr;
    int *ks = nullptr, *row_slot = nullptr, *row_jet = nullptr, *seg_beg = nullptr, *seg_end = nullptr;
    AttnItem* items = nullptr;
    CUtensorMap tm_act, tm_attn, tm_hidden, tm_q, tm_k, tm_k_ld, tm_vt, tm_resid;
    // staging for the host-buffer entry point
    uint8_t* stage = nullptr;
    size_t stage_bytes = 0;
};

struct Plan {
    int B = 0, D = 0, rows = 0;
    std::vector<int> row_slot, row_jet, seg_beg, seg_end;
    std::vector<AttnItem> items;
};

}  // namespace
}  // namespace mmf

using namespace mmf;

// kernel classes for launch counting and the optional per-class CUDA-event profile (bench.py roofline)
enum KernelClass { KC_PACK = 0, KC_EMBED_X, KC_GEMM_EMBED, KC_EMBED_FINISH, KC_GEMM_QKV, KC_ATTENTION, KC_GEMM_PROJ,
                   KC_GEMM_FC, KC_GEMM_MLP_OUT, KC_ADD_LN, KC_GEMM_HEAD, KC_HEAD_OUT_STEP, KC_UNPACK, KC_COUNT };
static const char* const kKernelClassNames[KC_COUNT] = {
    "pack", "embed_x", "gemm_embed", "embed_finish", "gemm_qkv", "attention", "gemm_attn_proj_resln", "gemm_mlp_fc_gelu",
    "gemm_mlp_out_resln", "add_layernorm", "gemm_head_fc_gelu", "head_out_step", "unpack"};

struct ProfRecord { int cls; cudaEvent_t e0, e1; double flops; };

struct MmfModel {
    MmfModelDesc desc{};
    int device = 0;
    DeviceArena arena;
    // transformer parameters
    std::vector<BlockDev> stream_blocks, main_blocks;   // ParticleFormer: x|y streams then fuse; Fused: main only
    const float *w0 = nullptr, *b0 = nullptr, *wxe2_b = nullptr, *ytab = nullptr, *ln1x_g = nullptr, *ln1x_b = nullptr;
    const float *lnmid_g = nullptr, *lnmid_b = nullptr, *lnfin_g = nullptr, *lnfin_b = nullptr;
    const bf16 *wxe2 = nullptr, *whead = nullptr;
    const float *bhead = nullptr, *wx2 = nullptr, *bx2 = nullptr, *wy2 = nullptr, *by2 = nullptr;
    CUtensorMap tm_wxe2, tm_whead;
    std::vector<float> time_expand_w, time_expand_b;    // host fp32, ParticleFormer only
    Workspace ws;
    EpicModel* epic = nullptr;          // EPiC has its own packed checkpoint, workspace and kernel
    TfTileModel* tile = nullptr;        // persistent per-tile kernel for jets of <= 128 particles (null: outside its envelope)
    int* d_err = nullptr;
    // side stream on which the persistent tile kernel runs while the layered path works on the jets of more than 128
    // particles (created on first use; fork / join with events around it)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t launches = 0;
    bool prof_on = false;
    std::vector<ProfRecord> prof_records;
    double prof_ms[KC_COUNT] = {0}, prof_flops[KC_COUNT] = {0};
    int64_t prof_count[KC_COUNT] = {0};
    double plan_sum_n2 = 0;          // sum over jets of n^2 (attention FLOP accounting)
    std::vector<float> h_temb, h_temb2;
    ~MmfModel() {
        arena.release();
        if (ws.base) cudaFree(ws.base);
        if (ws.stage) cudaFree(ws.stage);
        if (d_err) cudaFree(d_err);
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (epic) epic_destroy(epic);
        if (tile) tftile_destroy(tile);
    }
};

namespace mmf {
namespace {

#define MMF_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != 0) return _rc; \
    } while (0)

// Counts a launch and, when profiling is on, brackets it with CUDA events on the launching stream.
struct LaunchScope {
    MmfModel* m; int cls; cudaStream_t s; double flops; cudaEvent_t e0 = nullptr;
    LaunchScope(MmfModel* m_, int cls_, cudaStream_t s_, double flops_ = 0) : m(m_), cls(cls_), s(s_), flops(flops_) {
        m->launches += 1;
        if (m->prof_on) { cudaEventCreate(&e0); cudaEventRecord(e0, s); }
    }
    ~LaunchScope() {
        if (e0) {
            cudaEvent_t e1; cudaEventCreate(&e1); cudaEventRecord(e1, s);
            m->prof_records.push_back(ProfRecord{cls, e0, e1, flops});
        }
    }
};
#define MMF_LAUNCH(cls, flops, expr)                    \
    do {                                                \
        LaunchScope _scope(m, cls, c.stream, flops);    \
        int _rc = (expr);                               \
        if (_rc != 0) return _rc;                       \
    } while (0)

// ---------------------------------------------------------------------------------------------
// checkpoint packing
// ---------------------------------------------------------------------------------------------
// Stack `groups` sibling blocks (ParticleFormer's blocks_x.i / blocks_y.i) into one grouped parameter set.
BlockOff pack_block(DeviceArena& ar, WeightMap& wm, const std::vector<std::string>& prefixes, int C, int I, int H) {
    BlockOff o;
    o.groups = static_cast<int>(prefixes.size());
    o.C = C;
    const int hs = C / H;
    std::vector<float> wqkv, wproj, wfc, wp2, bqkv, bproj, bfc, bp2, ln1g, ln1b, ln2g, ln2b, qg, qb, kg, kb;
    o.qkln = wm.has(prefixes[0] + ".attn.q_layernorm.weight");
    for (const std::string& p : prefixes) {
        append(wqkv, wm.get(p + ".attn.c_attn.weight", 3 * C, C));
        append(bqkv, wm.get(p + ".attn.c_attn.bias", 3 * C, -1, true));
        append(wproj, wm.get(p + ".attn.c_proj.weight", C, C));
        append(bproj, wm.get(p + ".attn.c_proj.bias", C, -1, true));
        append(wfc, wm.get(p + ".ffw.c_fc.weight", I, C));
        append(bfc, wm.get(p + ".ffw.c_fc.bias", I, -1, true));
        append(wp2, wm.get(p + ".ffw.c_proj.weight", C, I));
        append(bp2, wm.get(p + ".ffw.c_proj.bias", C, -1, true));
        append(ln1g, wm.get(p + ".ln1.weight", C));
        append(ln1b, wm.get(p + ".ln1.bias", C, -1, true));
        append(ln2g, wm.get(p + ".ln2.weight", C));
        append(ln2b, wm.get(p + ".ln2.bias", C, -1, true));
        if (o.qkln) {
            append(qg, wm.get(p + ".attn.q_layernorm.weight", hs));
            append(qb, wm.get(p + ".attn.q_layernorm.bias", hs, -1, true));
            append(kg, wm.get(p + ".attn.k_layernorm.weight", hs));
            append(kb, wm.get(p + ".attn.k_layernorm.bias", hs, -1, true));
        }
    }
    o.wqkv = ar.put_bf16(wqkv); o.wproj = ar.put_bf16(wproj); o.wfc = ar.put_bf16(wfc); o.wp2 = ar.put_bf16(wp2);
    o.bqkv = ar.put_f32(bqkv); o.bproj = ar.put_f32(bproj); o.bfc = ar.put_f32(bfc); o.bp2 = ar.put_f32(bp2);
    o.ln1g = ar.put_f32(ln1g); o.ln1b = ar.put_f32(ln1b); o.ln2g = ar.put_f32(ln2g); o.ln2b = ar.put_f32(ln2b);
    if (o.qkln) { o.qg = ar.put_f32(qg); o.qb = ar.put_f32(qb); o.kg = ar.put_f32(kg); o.kb = ar.put_f32(kb); }
    return o;
}

int finish_block(const DeviceArena& ar, const BlockOff& o, int I, int H, BlockDev* d) {
    d->groups = o.groups; d->C = o.C; d->hs = o.C / H;
    d->wqkv = ar.at<bf16>(o.wqkv); d->wproj = ar.at<bf16>(o.wproj); d->wfc = ar.at<bf16>(o.wfc); d->wp2 = ar.at<bf16>(o.wp2);
    d->bqkv = ar.at<float>(o.bqkv); d->bproj = ar.at<float>(o.bproj); d->bfc = ar.at<float>(o.bfc); d->bp2 = ar.at<float>(o.bp2);
    d->ln1g = ar.at<float>(o.ln1g); d->ln1b = ar.at<float>(o.ln1b); d->ln2g = ar.at<float>(o.ln2g); d->ln2b = ar.at<float>(o.ln2b);
    d->qg = o.qkln ? ar.at<float>(o.qg) : nullptr; d->qb = o.qkln ? ar.at<float>(o.qb) : nullptr;
    d->kg = o.qkln ? ar.at<float>(o.kg) : nullptr; d->kb = o.qkln ? ar.at<float>(o.kb) : nullptr;
    const int G = o.groups, C = o.C;
    MMF_TRY(make_tmap_2d(&d->tm_wqkv, d->wqkv, 2, static_cast<uint64_t>(G) * 3 * C, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&d->tm_wproj, d->wproj, 2, static_cast<uint64_t>(G) * C, C, C, 64, C));
    MMF_TRY(make_tmap_2d(&d->tm_wfc, d->wfc, 2, static_cast<uint64_t>(G) * I, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&d->tm_wp2, d->wp2, 2, static_cast<uint64_t>(G) * C, I, I, 64, C));
    return 0;
}

float gelu_host(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

void layernorm_host(float* v, int n, const float* g, const float* b) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += v[i];
    const double mean = s / n;
    double q = 0;
    for (int i = 0; i < n; ++i) q += (v[i] - mean) * (v[i] - mean);
    const double rstd = 1.0 / std::sqrt(q / n + 1e-5);
    for (int i = 0; i < n; ++i) v[i] = static_cast<float>((v[i] - mean) * rstd) * g[i] + (b ? b[i] : 0.f);
}

int build_transformer(MmfModel* m, WeightMap& wm) {
    const MmfModelDesc& d = m->desc;
    const int E = d.n_embd, h = E / 2, I = d.n_inner, H = d.n_head, V = d.vocab_size;
    const std::string t = "transformer.";
    DeviceArena& ar = m->arena;

    const size_t o_w0 = ar.put_f32(wm.get(t + "wxe.0.weight", E, 3));
    const size_t o_b0 = ar.put_f32(wm.get(t + "wxe.0.bias", E));
    const size_t o_wxe2 = ar.put_bf16(wm.get(t + "wxe.2.weight", h, E));
    const size_t o_bxe2 = ar.put_f32(wm.get(t + "wxe.2.bias", h));
    const size_t o_l1g = ar.put_f32(wm.get(t + "ln1_x.weight", h));
    const size_t o_l1b = ar.put_f32(wm.get(t + "ln1_x.bias", h));

    // The whole discrete embedding branch is a function of the token only: fold it into a V x 128 table
    // ytab[k] = LN_ln1y( wye.2( GELU( wye.0[k] ) ) )   (reference ParticleTransformers.py:95-96, 190-191)
    std::vector<float> ytab(static_cast<size_t>(V) * h);
    {
        const std::vector<float> emb = wm.get(t + "wye.0.weight", V, E), w2 = wm.get(t + "wye.2.weight", h, E),
                                 b2 = wm.get(t + "wye.2.bias", h), g = wm.get(t + "ln1_y.weight", h),
                                 b = wm.get(t + "ln1_y.bias", h);
        for (int k = 0; k < V; ++k) {
            float* row = &ytab[static_cast<size_t>(k) * h];
            for (int o = 0; o < h; ++o) {
                double acc = b2[o];
                for (int i = 0; i < E; ++i) acc += static_cast<double>(gelu_host(emb[static_cast<size_t>(k) * E + i])) * w2[static_cast<size_t>(o) * E + i];
                row[o] = static_cast<float>(acc);
            }
            layernorm_host(row, h, g.data(), b.data());
        }
    }
    const size_t o_ytab = ar.put_f32(ytab);

    std::vector<BlockOff> stream_off, main_off;
    size_t o_mid_g = 0, o_mid_b = 0, o_fin_g = 0, o_fin_b = 0;
    if (d.arch == MMF_ARCH_PARTICLEFORMER) {
        for (int i = 0; i < d.n_layer; ++i)
            stream_off.push_back(pack_block(ar, wm, {t + "blocks_x." + std::to_string(i), t + "blocks_y." + std::to_string(i)}, h, I, H));
        for (int i = 0; i < d.n_layer_fused; ++i)
            main_off.push_back(pack_block(ar, wm, {t + "blocks_fuse." + std::to_string(i)}, E, I, H));
        std::vector<float> g = wm.get(t + "ln2_x.weight", h), b = wm.get(t + "ln2_x.bias", h);
        append(g, wm.get(t + "ln2_y.weight", h)); append(b, wm.get(t + "ln2_y.bias", h));
        o_mid_g = ar.put_f32(g); o_mid_b = ar.put_f32(b);
        g = wm.get(t + "ln3_x.weight", h); b = wm.get(t + "ln3_x.bias", h);
        append(g, wm.get(t + "ln3_y.weight", h)); append(b, wm.get(t + "ln3_y.bias", h));
        o_fin_g = ar.put_f32(g); o_fin_b = ar.put_f32(b);
        m->time_expand_w = wm.get(t + "time_expand.weight", E, h);
        m->time_expand_b = wm.get(t + "time_expand.bias", E);
    } else {
        for (int i = 0; i < d.n_layer; ++i)
            main_off.push_back(pack_block(ar, wm, {t + "blocks." + std::to_string(i)}, E, I, H));
        o_fin_g = ar.put_f32(wm.get(t + "ln2.weight", E));
        o_fin_b = ar.put_f32(wm.get(t + "ln2.bias", E));
    }
    std::vector<float> wh = wm.get(t + "head_x.0.weight", I, h), bh = wm.get(t + "head_x.0.bias", I);
    append(wh, wm.get(t + "head_y.0.weight", I, h)); append(bh, wm.get(t + "head_y.0.bias", I));
    const size_t o_wh = ar.put_bf16(wh), o_bh = ar.put_f32(bh);
    const size_t o_wx2 = ar.put_f32(wm.get(t + "head_x.2.weight", 3, I)), o_bx2 = ar.put_f32(wm.get(t + "head_x.2.bias", 3));
    const size_t o_wy2 = ar.put_f32(wm.get(t + "head_y.2.weight", V, I)), o_by2 = ar.put_f32(wm.get(t + "head_y.2.bias", V));
    if (!wm.missing.empty()) { set_last_error(wm.missing); return 2; }

    MMF_TRY(ar.upload());
    m->w0 = ar.at<float>(o_w0); m->b0 = ar.at<float>(o_b0); m->wxe2 = ar.at<bf16>(o_wxe2); m->wxe2_b = ar.at<float>(o_bxe2);
    m->ln1x_g = ar.at<float>(o_l1g); m->ln1x_b = ar.at<float>(o_l1b); m->ytab = ar.at<float>(o_ytab);
    if (d.arch == MMF_ARCH_PARTICLEFORMER) { m->lnmid_g = ar.at<float>(o_mid_g); m->lnmid_b = ar.at<float>(o_mid_b); }
    m->lnfin_g = ar.at<float>(o_fin_g); m->lnfin_b = ar.at<float>(o_fin_b);
    m->whead = ar.at<bf16>(o_wh); m->bhead = ar.at<float>(o_bh);
    m->wx2 = ar.at<float>(o_wx2); m->bx2 = ar.at<float>(o_bx2); m->wy2 = ar.at<float>(o_wy2); m->by2 = ar.at<float>(o_by2);
    m->stream_blocks.resize(stream_off.size());
    m->main_blocks.resize(main_off.size());
    for (size_t i = 0; i < stream_off.size(); ++i) MMF_TRY(finish_block(ar, stream_off[i], I, H, &m->stream_blocks[i]));
    for (size_t i = 0; i < main_off.size(); ++i) MMF_TRY(finish_block(ar, main_off[i], I, H, &m->main_blocks[i]));
    MMF_TRY(make_tmap_2d(&m->tm_wxe2, m->wxe2, 2, h, E, E, 64, 128));
    MMF_TRY(make_tmap_2d(&m->tm_whead, m->whead, 2, 2 * I, h, h, 64, 128));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
int ensure_workspace(MmfModel* m, int rows, int slots, int trows, int n_items) {
    Workspace& w = m->ws;
    if (rows <= w.mcap && slots <= w.slot_cap && trows <= w.tcap && n_items <= w.item_cap) return 0;
    // 25 % headroom on the row / item capacities (the packed row count changes from batch to batch)
    const int mcap = rows > w.mcap ? round_up(std::max(rows + rows / 4, 128), 128) : w.mcap;
    const int scap = std::max(w.slot_cap, slots);
    const int tcap = std::max(w.tcap, trows);
    const int icap = n_items > w.item_cap ? std::max(n_items + n_items / 4, 16) : w.item_cap;
    if (w.base) { MMF_CUDA_OK(cudaDeviceSynchronize()); MMF_CUDA_OK(cudaFree(w.base)); w.base = nullptr; }
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
    const size_t M = static_cast<size_t>(mcap);
    const size_t o_resid = take(M * 256 * 4), o_skip = take(M * 256 * 4), o_act = take(M * 256 * 2), o_q = take(M * 256 * 2),
                 o_k = take(M * 256 * 2), o_vt = take(M * 256 * 2), o_attn = take(M * 256 * 2), o_hid = take(M * 1024 * 2),
                 o_xs = take(M * 3 * 4), o_ks = take(M * 4), o_rs = take(M * 4), o_rj = take(M * 4), o_sb = take(M * 4),
                 o_se = take(M * 4), o_items = take(static_cast<size_t>(icap) * sizeof(AttnItem)),
                 o_temb = take(static_cast<size_t>(tcap) * 256 * 4), o_temb2 = take(static_cast<size_t>(tcap) * 256 * 4);
    MMF_CUDA_OK(cudaMalloc(&w.base, off));
    MMF_CUDA_OK(cudaMemset(w.base, 0, off));      // pad rows must stay finite: they are read as masked keys
    w.mcap = mcap; w.slot_cap = scap; w.tcap = tcap; w.item_cap = icap;
    w.resid = reinterpret_cast<float*>(w.base + o_resid); w.skip = reinterpret_cast<float*>(w.base + o_skip);
    w.act = reinterpret_cast<bf16*>(w.base + o_act); w.q = reinterpret_cast<bf16*>(w.base + o_q);
    w.k = reinterpret_cast<bf16*>(w.base + o_k); w.vt = reinterpret_cast<bf16*>(w.base + o_vt);
    w.attn = reinterpret_cast<bf16*>(w.base + o_attn); w.hidden = reinterpret_cast<bf16*>(w.base + o_hid);
    w.xs = reinterpret_cast<float*>(w.base + o_xs); w.ks = reinterpret_cast<int*>(w.base + o_ks);
    w.row_slot = reinterpret_cast<int*>(w.base + o_rs); w.row_jet = reinterpret_cast<int*>(w.base + o_rj);
    w.seg_beg = reinterpret_cast<int*>(w.base + o_sb); w.seg_end = reinterpret_cast<int*>(w.base + o_se);
    w.items = reinterpret_cast<AttnItem*>(w.base + o_items);
    w.temb = reinterpret_cast<float*>(w.base + o_temb); w.temb2 = reinterpret_cast<float*>(w.base + o_temb2);
    MMF_TRY(make_tmap_2d(&w.tm_act, w.act, 2, M, 256, 256, 64, 128));
    MMF_TRY(make_tmap_2d(&w.tm_attn, w.attn, 2, M, 256, 256, 64, 128));
    MMF_TRY(make_tmap_2d(&w.tm_hidden, w.hidden, 2, M, 1024, 1024, 64, 128));
    MMF_TRY(make_tmap_2d(&w.tm_q, w.q, 2, M, 256, 256, 64, 128));
    MMF_TRY(make_tmap_2d(&w.tm_k, w.k, 2, M, 256, 256, 64, 128));
    MMF_TRY(make_tmap_2d(&w.tm_k_ld, w.k, 2, M, 256, 256, 64, 32));
    MMF_TRY(make_tmap_2d(&w.tm_vt, w.vt, 2, 256, M, M, 64, 64));
    MMF_TRY(make_tmap_2d(&w.tm_resid, w.resid, 4, M, 256, 256, 32, 128));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// planning: packed row layout + attention work items from the (host) mask
// ---------------------------------------------------------------------------------------------
// TMA needs the innermost coordinate (the token index of V^T) on a 16-byte boundary, so the key window of an
// item starts at the jet's first row rounded down to a multiple of 8; the few foreign rows in front are masked
// by the per-row segment bounds like any other jet in the window.
void plan_attention_items(const std::vector<int>& jet_start, const std::vector<int>& jet_n, std::vector<AttnItem>* items) {
    int g_start = -1, g_rows = 0;
    auto window = [](int q_row0, int nq, int first_key, int end_key) {
        const int k0 = first_key & ~7;
        return AttnItem{q_row0, nq, k0, end_key - k0};
    };
    auto flush = [&]() {
        if (g_rows > 0) items->push_back(window(g_start, g_rows, g_start, g_start + g_rows));
        g_start = -1; g_rows = 0;
    };
    for (size_t j = 0; j < jet_n.size(); ++j) {
        const int n = jet_n[j], s = jet_start[j];
        if (n == 0) continue;
        if (n > kTileM) {
            flush();
            items->push_back(window(s, kTileM, s, s + n));
            items->push_back(window(s + kTileM, n - kTileM, s, s + n));
            continue;
        }
        if (g_rows + n > kTileM) flush();
        if (g_rows == 0) g_start = s;
        g_rows += n;
    }
    flush();
}

int build_plan(const int64_t* mask, int B, int D, Plan* p) {
    p->B = B; p->D = D;
    p->row_slot.clear(); p->row_jet.clear(); p->seg_beg.clear(); p->seg_end.clear(); p->items.clear();
    std::vector<int> jet_start(B), jet_n(B);
    int rows = 0;
    for (int b = 0; b < B; ++b) {
        jet_start[b] = rows;
        for (int d = 0; d < D; ++d)
            if (mask[static_cast<size_t>(b) * D + d] != 0) { p->row_slot.push_back(b * D + d); p->row_jet.push_back(b); ++rows; }
        jet_n[b] = rows - jet_start[b];
        MMF_REQUIRE(jet_n[b] <= kMaxKeys - 8, "a jet has more than 152 real particles");
        for (int i = 0; i < jet_n[b]; ++i) { p->seg_beg.push_back(jet_start[b]); p->seg_end.push_back(rows); }
    }
    p->rows = rows;
    plan_attention_items(jet_start, jet_n, &p->items);
    return 0;
}

int upload_plan(MmfModel* m, const Plan& p, cudaStream_t s) {
    Workspace& w = m->ws;
    m->plan_sum_n2 = 0;
    for (size_t r = 0; r < p.seg_beg.size(); ++r) m->plan_sum_n2 += p.seg_end[r] - p.seg_beg[r];   // sum_rows n_jet = sum_jets n^2
    const size_t n = static_cast<size_t>(p.rows);
    if (n) {
        MMF_CUDA_OK(cudaMemcpyAsync(w.row_slot, p.row_slot.data(), n * 4, cudaMemcpyHostToDevice, s));
        MMF_CUDA_OK(cudaMemcpyAsync(w.row_jet, p.row_jet.data(), n * 4, cudaMemcpyHostToDevice, s));
        MMF_CUDA_OK(cudaMemcpyAsync(w.seg_beg, p.seg_beg.data(), n * 4, cudaMemcpyHostToDevice, s));
        MMF_CUDA_OK(cudaMemcpyAsync(w.seg_end, p.seg_end.data(), n * 4, cudaMemcpyHostToDevice, s));
    }
    // rows of the last partial tile: keep maps in range (jet 0, empty segment)
    const int padded = round_up(std::max(p.rows, 1), kTileM);
    if (padded > p.rows) {
        const size_t tail = static_cast<size_t>(padded - p.rows) * 4;
        MMF_CUDA_OK(cudaMemsetAsync(w.row_slot + n, 0, tail, s));
        MMF_CUDA_OK(cudaMemsetAsync(w.row_jet + n, 0, tail, s));
        MMF_CUDA_OK(cudaMemsetAsync(w.seg_beg + n, 0, tail, s));
        MMF_CUDA_OK(cudaMemsetAsync(w.seg_end + n, 0, tail, s));
    }
    if (!p.items.empty())
        MMF_CUDA_OK(cudaMemcpyAsync(w.items, p.items.data(), p.items.size() * sizeof(AttnItem), cudaMemcpyHostToDevice, s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// time tables: sin/cos features (reference utils/models.py:62-75) and time_expand (ParticleTransformers.py:109)
// ---------------------------------------------------------------------------------------------
int upload_time_tables(MmfModel* m, const float* times, int n, cudaStream_t s) {
    const MmfModelDesc& d = m->desc;
    m->h_temb.assign(static_cast<size_t>(n) * 256, 0.f);
    const bool pf = d.arch == MMF_ARCH_PARTICLEFORMER;
    if (pf) m->h_temb2.assign(static_cast<size_t>(n) * 256, 0.f);
    for (int i = 0; i < n; ++i) {
        float* row = &m->h_temb[static_cast<size_t>(i) * 256];
        if (pf) {
            sincos_row(times[i], 128, row);
            memcpy(row + 128, row, 128 * sizeof(float));        // same embedding for the x and y streams
            float* r2 = &m->h_temb2[static_cast<size_t>(i) * 256];
            for (int o = 0; o < 256; ++o) {
                float acc = 0.f;
                const float* wrow = &m->time_expand_w[static_cast<size_t>(o) * 128];
                for (int j = 0; j < 128; ++j) acc += wrow[j] * row[j];
                r2[o] = acc + m->time_expand_b[o];
            }
        } else {
            sincos_row(times[i], 256, row);
        }
    }
    MMF_CUDA_OK(cudaMemcpyAsync(m->ws.temb, m->h_temb.data(), m->h_temb.size() * 4, cudaMemcpyHostToDevice, s));
    if (pf) MMF_CUDA_OK(cudaMemcpyAsync(m->ws.temb2, m->h_temb2.data(), m->h_temb2.size() * 4, cudaMemcpyHostToDevice, s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// one encoder forward over the packed rows (+ fused output projection / step)
// ---------------------------------------------------------------------------------------------
struct ForwardCtx {
    int rows, n_items;
    const float* temb;       // [*,256] first row to use
    const float* temb2;      // ParticleFormer fuse-stream embedding
    const int* row_jet;      // null when time is uniform over the batch
    HeadOutArgs head;        // destination of the fused tail
    cudaStream_t stream;
};

int run_block(MmfModel* m, const BlockDev& b, const ForwardCtx& c, const float* temb, const float* next_g, const float* next_b) {
    Workspace& w = m->ws;
    const int m_tiles = round_up(std::max(c.rows, 1), kTileM) / kTileM;
    const int G = b.groups, C = b.C, I = m->desc.n_inner;
    GemmArgs a{};
    // fused QKV projection, per-head LayerNorm on q and k, V stored transposed
    a.kblocks = C / kBK; a.a_col_group_stride = C; a.w_rows_per_group = 3 * C; a.bias = b.bqkv;
    a.sect_width = C; a.hs = b.hs; a.q_g = b.qg; a.q_b = b.qb; a.k_g = b.kg; a.k_b = b.kb; a.vt = w.vt; a.vt_ld = w.mcap;
    const double rows = c.rows;
    MMF_LAUNCH(KC_GEMM_QKV, 2.0 * rows * C * 3 * C * G, launch_gemm(EPI_QKV, 128, w.tm_act, b.tm_wqkv, w.tm_q, w.tm_k, a, m_tiles, 3 * C / 128, G, c.stream));
    // masked attention
    AttnArgs at{};
    at.items = w.items; at.seg_beg = w.seg_beg; at.seg_end = w.seg_end; at.out = w.attn; at.ld_out = 256; at.hs = b.hs;
    at.scale_log2e = 1.4426950408889634f / std::sqrt(static_cast<float>(b.hs));
    MMF_LAUNCH(KC_ATTENTION, 4.0 * m->plan_sum_n2 * 256, launch_attention(w.tm_q, w.tm_k_ld, w.tm_vt, at, c.n_items, 4, c.stream));
    // attention projection + residual, LayerNorm ln2 -> MLP operand
    a = GemmArgs{};
    a.kblocks = C / kBK; a.a_col_group_stride = C; a.w_rows_per_group = C; a.bias = b.bproj;
    a.ln_g = b.ln2g; a.ln_b = b.ln2b;
    MMF_LAUNCH(KC_GEMM_PROJ, 2.0 * rows * C * C * G, launch_gemm(EPI_RESLN, C, w.tm_attn, b.tm_wproj, w.tm_act, w.tm_resid, a, m_tiles, 1, G, c.stream));
    // MLP up-projection + exact GELU
    a = GemmArgs{};
    a.kblocks = C / kBK; a.a_col_group_stride = C; a.w_rows_per_group = I; a.bias = b.bfc; a.act = 1; a.out_col_group_stride = I;
    MMF_LAUNCH(KC_GEMM_FC, 2.0 * rows * C * I * G, launch_gemm(EPI_STORE_BF16, 128, w.tm_act, b.tm_wfc, w.tm_hidden, w.tm_hidden, a, m_tiles, I / 128, G, c.stream));
    // MLP down-projection + residual + time embedding, LayerNorm of the next block -> its QKV operand
    a = GemmArgs{};
    a.kblocks = I / kBK; a.a_col_group_stride = I; a.w_rows_per_group = C; a.bias = b.bp2;
    a.temb = temb; a.temb_ld = 256; a.row_jet = c.row_jet; a.ln_g = next_g; a.ln_b = next_b;
    MMF_LAUNCH(KC_GEMM_MLP_OUT, 2.0 * rows * C * I * G, launch_gemm(EPI_RESLN, C, w.tm_hidden, b.tm_wp2, w.tm_act, w.tm_resid, a, m_tiles, 1, G, c.stream));
    return 0;
}

int run_forward(MmfModel* m, const ForwardCtx& c) {
    Workspace& w = m->ws;
    const MmfModelDesc& d = m->desc;
    const int m_tiles = round_up(std::max(c.rows, 1), kTileM) / kTileM;
    const int rows_padded = m_tiles * kTileM;
    const bool pf = d.arch == MMF_ARCH_PARTICLEFORMER;
    // wxe: Linear(3,256) + GELU on CUDA cores, Linear(256,128) on tensor cores (raw output into resid[:, :128])
    const double rows = c.rows;
    MMF_LAUNCH(KC_EMBED_X, 2.0 * rows * 3 * d.n_embd, launch_embed_x(w.xs, c.rows, m->w0, m->b0, d.n_embd, 1, w.hidden, 1024, c.stream));
    GemmArgs a{};
    a.kblocks = d.n_embd / kBK; a.w_rows_per_group = 128; a.bias = m->wxe2_b;
    MMF_LAUNCH(KC_GEMM_EMBED, 2.0 * rows * d.n_embd * 128, launch_gemm(EPI_STORE_F32, 128, w.tm_hidden, m->tm_wxe2, w.tm_resid, w.tm_resid, a, m_tiles, 1, 1, c.stream));
    const std::vector<BlockDev>& first = pf ? m->stream_blocks : m->main_blocks;
    EmbedFinishArgs ef{};
    ef.rows = rows_padded; ef.resid = w.resid; ef.skip = w.skip; ef.act = w.act; ef.ks = w.ks; ef.ytab = m->ytab;
    ef.ln1x_g = m->ln1x_g; ef.ln1x_b = m->ln1x_b; ef.temb = c.temb; ef.temb_ld = 256; ef.row_jet = c.row_jet;
    ef.next_ln_width = first[0].C; ef.next_g = first[0].ln1g; ef.next_b = first[0].ln1b;
    MMF_LAUNCH(KC_EMBED_FINISH, 0, launch_embed_finish(ef, c.stream));

    if (pf) {
        for (size_t i = 0; i < m->stream_blocks.size(); ++i) {
            const bool last = i + 1 == m->stream_blocks.size();
            MMF_TRY(run_block(m, m->stream_blocks[i], c, c.temb, last ? nullptr : m->stream_blocks[i + 1].ln1g,
                              last ? nullptr : m->stream_blocks[i + 1].ln1b));
        }
        AddLnArgs j{};                       // x = ln2_x(x + x_skip) | y = ln2_y(y + y_skip); z = cat + temb2
        j.rows = rows_padded; j.resid = w.resid; j.skip = w.skip; j.act = w.act; j.ln1_width = 128; j.ln1_g = m->lnmid_g;
        j.ln1_b = m->lnmid_b; j.temb = c.temb2; j.temb_ld = 256; j.row_jet = c.row_jet; j.write_resid = 1;
        j.ln2_width = 256; j.ln2_g = m->main_blocks[0].ln1g; j.ln2_b = m->main_blocks[0].ln1b;
        MMF_LAUNCH(KC_ADD_LN, 0, launch_add_ln(j, c.stream));
    }
    const float* main_temb = pf ? c.temb2 : c.temb;
    for (size_t i = 0; i < m->main_blocks.size(); ++i) {
        const bool last = i + 1 == m->main_blocks.size();
        MMF_TRY(run_block(m, m->main_blocks[i], c, main_temb, last ? nullptr : m->main_blocks[i + 1].ln1g,
                          last ? nullptr : m->main_blocks[i + 1].ln1b));
    }
    AddLnArgs f{};                           // ParticleFormer: ln3_x | ln3_y on (z + skip); Fused: ln2 over 256
    f.rows = rows_padded; f.resid = w.resid; f.skip = w.skip; f.act = w.act; f.ln1_width = pf ? 128 : 256;
    f.ln1_g = m->lnfin_g; f.ln1_b = m->lnfin_b;
    MMF_LAUNCH(KC_ADD_LN, 0, launch_add_ln(f, c.stream));
    // heads: Linear(128,512)+GELU for both heads as one 2-group GEMM, then the tiny projections fused with the step
    a = GemmArgs{};
    a.kblocks = 128 / kBK; a.a_col_group_stride = 128; a.w_rows_per_group = d.n_inner; a.bias = m->bhead; a.act = 1;
    a.out_col_group_stride = d.n_inner;
    MMF_LAUNCH(KC_GEMM_HEAD, 2.0 * rows * 128 * d.n_inner * 2, launch_gemm(EPI_STORE_BF16, 128, w.tm_act, m->tm_whead, w.tm_hidden, w.tm_hidden, a, m_tiles, d.n_inner / 128, 2, c.stream));
    HeadOutArgs ho = c.head;
    ho.rows = c.rows; ho.hidden = w.hidden; ho.ld_hidden = 1024; ho.wx = m->wx2; ho.bx = m->bx2; ho.wy = m->wy2; ho.by = m->by2;
    ho.row_slot = w.row_slot; ho.xs = w.xs; ho.ks = w.ks;
    MMF_LAUNCH(KC_HEAD_OUT_STEP, 2.0 * rows * d.n_inner * (3 + d.vocab_size), launch_head_out(ho, d.vocab_size, c.stream));
    return 0;
}

int check_device_flags(MmfModel* m, cudaStream_t s) {
    int flag = 0;
    MMF_CUDA_OK(cudaMemcpyAsync(&flag, m->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMF_CUDA_OK(cudaStreamSynchronize(s));
    if (flag) {
        MMF_CUDA_OK(cudaMemsetAsync(m->d_err, 0, sizeof(int), s));
        set_last_error("Values in `k` outside of bound [0, vocab_size)");      // reference model/MJB.py:177-182
        return 3;
    }
    return 0;
}

int fetch_mask(const int64_t* mask_dev, int B, int D, std::vector<int64_t>* host, cudaStream_t s) {
    host->resize(static_cast<size_t>(B) * D);
    MMF_CUDA_OK(cudaMemcpyAsync(host->data(), mask_dev, host->size() * 8, cudaMemcpyDeviceToHost, s));
    MMF_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int generate_device(MmfModel* m, const float* x0, const int64_t* k0, const int64_t* mask_host, int B, int D,
                    const float* t_grid, int N, float dt, const MmfStepOptions* opts, const float* u,
                    const uint8_t* forced_k, float* x_out, int64_t* k_out, float* rates_out, cudaStream_t s) {
    const MmfModelDesc& d = m->desc;
    MMF_REQUIRE(N >= 1 && B >= 1 && D >= 1 && D <= d.max_num_particles, "bad problem shape");
    if (d.arch == MMF_ARCH_EPIC) return epic_generate(m->epic, x0, mask_host, B, D, t_grid, N, dt, x_out, s);
    MMF_REQUIRE(opts != nullptr && k0 != nullptr && k_out != nullptr, "the transformers need tokens and step options");
    MMF_REQUIRE(opts->method == 0, "the sampler loop uses the tau-leap step (hard-coded in the reference, model/solvers.py:9)");
    const size_t slots = static_cast<size_t>(B) * D;
    // jets of <= 128 particles run in the persistent tile kernel; the layered path below takes what is left
    std::vector<int64_t> rest;
    const int64_t* lmask = mask_host;
    if (m->tile) {
        std::vector<unsigned char> handled;
        TfRunArgs ta{};
        ta.x0 = x0; ta.k0 = reinterpret_cast<const long long*>(k0); ta.mask_host = mask_host; ta.B = B; ta.D = D;
        ta.times = t_grid; ta.n_times = N; ta.per_jet_time = false; ta.nsteps = N; ta.dt = dt; ta.opts = opts; ta.u = u;
        ta.forced = forced_k; ta.x_out = x_out; ta.k_out = reinterpret_cast<long long*>(k_out); ta.rates_out = rates_out;
        ta.err_flag = m->d_err;
        MMF_TRY(tftile_prepare(m->tile, ta, &handled, s));
        rest.assign(mask_host, mask_host + slots);
        for (int b = 0; b < B; ++b)
            if (handled[b]) std::fill(rest.begin() + static_cast<size_t>(b) * D, rest.begin() + static_cast<size_t>(b + 1) * D, 0);
        lmask = rest.data();
    }
    Plan plan;
    MMF_TRY(build_plan(lmask, B, D, &plan));
    Workspace& w = m->ws;
    if (plan.rows > 0) {
        MMF_TRY(ensure_workspace(m, plan.rows, B * D, N, static_cast<int>(plan.items.size())));
        MMF_TRY(upload_plan(m, plan, s));
        MMF_TRY(upload_time_tables(m, t_grid, N, s));
        LaunchScope sc(m, KC_PACK, s);
        MMF_TRY(launch_pack(x0, reinterpret_cast<const long long*>(k0), w.row_slot, plan.rows, d.vocab_size, w.xs, w.ks, m->d_err, s));
    }
    // both packed states are taken: from here on the outputs may alias the inputs.  Padded slots read as zero.
    MMF_CUDA_OK(cudaMemsetAsync(x_out, 0, slots * 3 * 4, s));
    MMF_CUDA_OK(cudaMemsetAsync(k_out, 0, slots * 8, s));
    if (rates_out) MMF_CUDA_OK(cudaMemsetAsync(rates_out, 0, slots * d.vocab_size * 4, s));
    // A batch with both kinds of jets: the tile kernel goes to a side stream and the layered launches follow on `s`, so the
    // few jets of more than 128 particles no longer wait behind (nor delay) the tiles - they fill SMs the tiles leave free.
    // (Not under the per-class profile or the trace build, which time / read back on `s`.)
    bool forked = false;
    if (m->tile) {
        if (plan.rows > 0 && !m->prof_on && getenv("MMF_TRACE") == nullptr && getenv("MMF_NO_OVERLAP") == nullptr) {
            if (!m->side) {
                MMF_CUDA_OK(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
                MMF_CUDA_OK(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
                MMF_CUDA_OK(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
            }
            MMF_CUDA_OK(cudaEventRecord(m->ev_fork, s));
            MMF_CUDA_OK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
            MMF_TRY(tftile_launch(m->tile, m->side));
            MMF_CUDA_OK(cudaEventRecord(m->ev_join, m->side));
            forked = true;
        } else {
            MMF_TRY(tftile_launch(m->tile, s));
        }
    }
    if (plan.rows == 0) return 0;
    int rc = 0;
    for (int i = 0; i < N; ++i) {
        ForwardCtx c{};
        c.rows = plan.rows; c.n_items = static_cast<int>(plan.items.size());
        c.temb = w.temb + static_cast<size_t>(i) * 256; c.temb2 = w.temb2 + static_cast<size_t>(i) * 256; c.row_jet = nullptr;
        c.stream = s;
        HeadOutArgs& h = c.head;
        h.do_step = 1;
        h.sl.sp = StepParams{opts->temperature, dt, opts->beta, opts->top_p, opts->top_k, d.vocab_size};
        h.sl.u = u ? u + static_cast<size_t>(i) * slots * d.vocab_size : nullptr;
        h.sl.seed = opts->seed; h.sl.slot0 = opts->first_global_jet * static_cast<uint64_t>(D); h.sl.step = static_cast<uint32_t>(i);
        h.sl.err_flag = m->d_err;
        det_thermostat(t_grid[i], opts->beta, d.vocab_size, &h.w, &h.coef);
        h.forced = forced_k ? forced_k + static_cast<size_t>(i) * slots : nullptr;
        const bool last = i + 1 == N;
        h.rates_out = last ? rates_out : nullptr;
        h.argmax_out = (last && opts->use_final_max_rates) ? 1 : 0;
        rc = run_forward(m, c);
        if (rc != 0) break;
    }
    if (rc == 0) { LaunchScope sc(m, KC_UNPACK, s); rc = launch_unpack(w.xs, w.ks, w.row_slot, plan.rows, x_out, reinterpret_cast<long long*>(k_out), s); }
    if (forked) MMF_CUDA_OK(cudaStreamWaitEvent(s, m->ev_join, 0));      // joined on the error paths too: `s` never runs ahead of the tiles
    return rc;
}

}  // namespace
}  // namespace mmf

// =============================================================================================
// extern "C"
// =============================================================================================
extern "C" {

// out-of-range-token flag of the standalone step, one per device (the step has no model handle to hang it on)
static int* g_step_flag[64] = {nullptr};
static int step_flag(int device, int** out) {
    MMF_REQUIRE(device >= 0 && device < 64, "device index out of range");
    if (!g_step_flag[device]) {
        MMF_CUDA_OK(cudaMalloc(&g_step_flag[device], sizeof(int)));
        MMF_CUDA_OK(cudaMemset(g_step_flag[device], 0, sizeof(int)));
    }
    *out = g_step_flag[device];
    return 0;
}

int mmf_abi_version(void) { return MMF_ABI_VERSION; }
const char* mmf_last_error(void) { return g_last_error.c_str(); }

int mmf_model_create(const MmfModelDesc* desc, const MmfWeightRef* weights, int32_t n_weights, int32_t device, MmfModel** out) {
    MMF_REQUIRE(desc && weights && out, "null argument");
    *out = nullptr;
    MMF_REQUIRE(desc->arch >= 0 && desc->arch <= MMF_ARCH_EPIC, "unknown architecture");
    MMF_REQUIRE(desc->n_embd == 256 && desc->dim_continuous == 3 &&
                    (desc->arch == MMF_ARCH_EPIC || (desc->n_inner == 512 && desc->n_head == 4)),
                "accelerated path is built for n_embd=256, n_inner=512, n_head=4, dim_continuous=3");
    MMF_REQUIRE(desc->vocab_size >= 2 && desc->vocab_size <= kMaxV, "vocab_size must be in [2,16]");
    MMF_REQUIRE(desc->max_num_particles >= 1 && desc->max_num_particles <= kMaxKeys - 8, "max_num_particles must be <= 152");
    MMF_REQUIRE(desc->n_layer >= 1 && (desc->arch != MMF_ARCH_PARTICLEFORMER || desc->n_layer_fused >= 1), "need at least one block");
    int ndev = 0;
    MMF_CUDA_OK(cudaGetDeviceCount(&ndev));
    MMF_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
    MMF_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    MMF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    MMF_REQUIRE(prop.major == 10, "libmmf_b200 contains sm_100a code only; this device is not a Blackwell B200-class GPU");
    std::unique_ptr<MmfModel> m(new MmfModel());
    m->desc = *desc;
    m->device = device;
    WeightMap wm;
    for (int i = 0; i < n_weights; ++i) wm.m[weights[i].name] = &weights[i];
    int rc = desc->arch == MMF_ARCH_EPIC ? epic_create(*desc, wm, &m->epic) : build_transformer(m.get(), wm);
    if (rc) return rc;
    if (desc->arch != MMF_ARCH_EPIC && !getenv("MMF_NO_TILE_KERNEL")) {       // env switch: A/B runs of the layered path
        rc = tftile_create(*desc, wm, &m->tile);
        if (rc) return rc;
    }
    MMF_CUDA_OK(cudaMalloc(&m->d_err, sizeof(int)));
    MMF_CUDA_OK(cudaMemset(m->d_err, 0, sizeof(int)));
    *out = m.release();
    return 0;
}

void mmf_model_destroy(MmfModel* model) {
    if (!model) return;
    cudaSetDevice(model->device);
    cudaDeviceSynchronize();
    delete model;
}

int64_t mmf_launch_count(const MmfModel* model) {
    return model ? model->launches + epic_launches(model->epic) + tftile_launches(model->tile) : 0;
}

int mmf_profile_enable(MmfModel* m, int32_t on) {
    MMF_REQUIRE(m != nullptr, "null model");
    m->prof_on = on != 0;
    return 0;
}
int32_t mmf_profile_num_classes(void) { return KC_COUNT; }
const char* mmf_profile_class_name(int32_t i) { return (i >= 0 && i < KC_COUNT) ? kKernelClassNames[i] : ""; }
int mmf_profile_read(MmfModel* m, double* ms, int64_t* launches, double* flops, int32_t reset) {
    MMF_REQUIRE(m != nullptr, "null model");
    MMF_CUDA_OK(cudaSetDevice(m->device));
    MMF_CUDA_OK(cudaDeviceSynchronize());
    for (ProfRecord& r : m->prof_records) {
        float t = 0.f;
        MMF_CUDA_OK(cudaEventElapsedTime(&t, r.e0, r.e1));
        m->prof_ms[r.cls] += t; m->prof_count[r.cls] += 1; m->prof_flops[r.cls] += r.flops;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    m->prof_records.clear();
    for (int i = 0; i < KC_COUNT; ++i) {
        if (ms) ms[i] = m->prof_ms[i];
        if (launches) launches[i] = m->prof_count[i];
        if (flops) flops[i] = m->prof_flops[i];
        if (reset) { m->prof_ms[i] = 0; m->prof_count[i] = 0; m->prof_flops[i] = 0; }
    }
    return 0;
}

int mmf_encoder_forward(MmfModel* m, const float* x, const int64_t* k, const int64_t* mask, const float* t, int32_t B,
                        int32_t D, float* vt_out, float* logits_out, void* stream) {
    MMF_REQUIRE(m && x && mask && t && vt_out, "null argument");
    MMF_REQUIRE(B >= 1 && D >= 1 && D <= m->desc.max_num_particles, "bad batch shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_CUDA_OK(cudaSetDevice(m->device));
    std::vector<int64_t> hmask;
    MMF_TRY(fetch_mask(mask, B, D, &hmask, s));
    std::vector<float> ht(B);
    MMF_CUDA_OK(cudaMemcpyAsync(ht.data(), t, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, s));
    MMF_CUDA_OK(cudaStreamSynchronize(s));
    if (m->desc.arch == MMF_ARCH_EPIC) return epic_forward(m->epic, x, hmask.data(), ht.data(), B, D, vt_out, s);
    MMF_REQUIRE(k && logits_out, "the transformers need tokens and a logits buffer");
    const size_t slots = static_cast<size_t>(B) * D;
    MMF_CUDA_OK(cudaMemsetAsync(vt_out, 0, slots * 3 * 4, s));
    MMF_CUDA_OK(cudaMemsetAsync(logits_out, 0, slots * m->desc.vocab_size * 4, s));
    if (m->tile) {
        std::vector<unsigned char> handled;
        TfRunArgs ta{};
        ta.x0 = x; ta.k0 = reinterpret_cast<const long long*>(k); ta.mask_host = hmask.data(); ta.B = B; ta.D = D;
        ta.times = ht.data(); ta.n_times = B; ta.per_jet_time = true; ta.nsteps = 1; ta.vt_out = vt_out; ta.logits_out = logits_out;
        ta.err_flag = m->d_err;
        MMF_TRY(tftile_prepare(m->tile, ta, &handled, s));
        MMF_TRY(tftile_launch(m->tile, s));
        for (int b = 0; b < B; ++b)
            if (handled[b]) std::fill(hmask.begin() + static_cast<size_t>(b) * D, hmask.begin() + static_cast<size_t>(b + 1) * D, 0);
    }
    Plan plan;
    MMF_TRY(build_plan(hmask.data(), B, D, &plan));
    if (plan.rows > 0) {
        MMF_TRY(ensure_workspace(m, plan.rows, B * D, B, static_cast<int>(plan.items.size())));
        MMF_TRY(upload_plan(m, plan, s));
        MMF_TRY(upload_time_tables(m, ht.data(), B, s));
        Workspace& w = m->ws;
        { LaunchScope sc(m, KC_PACK, s); MMF_TRY(launch_pack(x, reinterpret_cast<const long long*>(k), w.row_slot, plan.rows, m->desc.vocab_size, w.xs, w.ks, m->d_err, s)); }
        ForwardCtx c{};
        c.rows = plan.rows; c.n_items = static_cast<int>(plan.items.size());
        c.temb = w.temb; c.temb2 = w.temb2; c.row_jet = w.row_jet; c.stream = s;
        c.head.vt_out = vt_out; c.head.logits_out = logits_out; c.head.do_step = 0;
        c.head.sl.sp.vocab = m->desc.vocab_size;
        MMF_TRY(run_forward(m, c));
    }
    return check_device_flags(m, s);
}

int mmf_hybrid_step_status(int32_t device, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_CUDA_OK(cudaSetDevice(device));
    int* d = nullptr;
    MMF_TRY(step_flag(device, &d));
    int flag = 0;
    MMF_CUDA_OK(cudaMemcpyAsync(&flag, d, sizeof(int), cudaMemcpyDeviceToHost, s));
    MMF_CUDA_OK(cudaStreamSynchronize(s));
    if (flag) {
        MMF_CUDA_OK(cudaMemsetAsync(d, 0, sizeof(int), s));
        set_last_error("Values in `k` outside of bound [0, vocab_size)");      // reference model/MJB.py:177-182
        return 3;
    }
    return 0;
}

int mmf_hybrid_step(const float* vt, const float* logits, float* x, int64_t* k, const float* t, float dt,
                    const MmfStepOptions* opts, const float* u, uint32_t step_index, int32_t B, int32_t D, int32_t V,
                    float* rates_out, int32_t device, void* stream) {
    MMF_REQUIRE(vt && logits && x && k && t && opts, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_CUDA_OK(cudaSetDevice(device));
    int* d_flag_dev = nullptr;
    MMF_TRY(step_flag(device, &d_flag_dev));
    MMF_REQUIRE(opts->method == 0 || opts->method == 1, "MmfStepOptions.method must be 0 (tau-leap) or 1 (categorical Euler)");
    MMF_REQUIRE(opts->method == 0 || opts->temperature == 1.0f,
                "the categorical Euler step is defined for temperature 1 only (reference solvers.py:95-99 rescales per class for a fixed batch shape)");
    StepLaunch sl{};
    sl.sp = StepParams{opts->temperature, dt, opts->beta, opts->top_p, opts->top_k, V, opts->method};
    sl.u = u; sl.seed = opts->seed; sl.slot0 = opts->first_global_jet * static_cast<uint64_t>(D); sl.step = step_index;
    sl.err_flag = d_flag_dev;
    return launch_hybrid_step(vt, logits, x, reinterpret_cast<long long*>(k), t, B, D, sl, rates_out, s);
}

int mmf_euler_step(const float* vt, float* x, float dt, int64_t n, int32_t device, void* stream) {
    MMF_REQUIRE(vt && x, "null argument");
    MMF_CUDA_OK(cudaSetDevice(device));
    return launch_euler(vt, x, dt, n, static_cast<cudaStream_t>(stream));
}

int mmf_jet_observables(const float* x, const int64_t* k, const int64_t* mask, const float* mean, const float* std_, int32_t B,
                        int32_t D, int32_t V, float* kin_out, int32_t* counts_out, int32_t device, void* stream) {
    MMF_REQUIRE(x && mask && kin_out, "null argument");
    MMF_REQUIRE(B >= 0 && D >= 1 && D <= 4080, "D must be in [1, 4080] (8-bit per-lane token counters, 16 lanes per jet)");
    MMF_REQUIRE(counts_out == nullptr || (k != nullptr && V >= 1 && V <= 16), "token counts need k and 1 <= V <= 16");
    MMF_CUDA_OK(cudaSetDevice(device));
    ObsArgs a{};
    a.B = B; a.D = D; a.V = V;
    for (int c = 0; c < 3; ++c) { a.mean[c] = mean ? mean[c] : 0.0f; a.std[c] = std_ ? std_[c] : 1.0f; }
    a.kin = kin_out;
    a.counts = counts_out;
    return launch_jet_observables(x, reinterpret_cast<const long long*>(k), reinterpret_cast<const long long*>(mask), a,
                                  static_cast<cudaStream_t>(stream));
}

int mmf_make_source(const float* mult_probs, int32_t B, int32_t D, int32_t V, uint64_t seed, uint64_t first_global_jet, float* x0,
                    int64_t* k0, int64_t* mask, int32_t* n_out, int32_t device, void* stream) {
    MMF_REQUIRE(B >= 0 && D >= 1 && D <= kSrcMaxD, "max_num_particles must be in [1, 255]");
    if (B == 0) return 0;                                 // an empty shard (more ranks than jets): nothing to draw
    MMF_REQUIRE(mult_probs && x0 && mask && n_out, "null argument");
    MMF_REQUIRE(k0 == nullptr || V >= 2, "vocab_size must be at least 2 when tokens are requested");
    double tot = 0.0;
    for (int m = 0; m <= D; ++m) {
        MMF_REQUIRE(mult_probs[m] >= 0.0f, "multiplicity probabilities must be non-negative");
        tot += static_cast<double>(mult_probs[m]);
    }
    MMF_REQUIRE(tot > 0.0, "multiplicity probabilities sum to zero");
    MMF_CUDA_OK(cudaSetDevice(device));
    SourceArgs a{};
    a.B = B; a.D = D; a.V = V; a.seed = seed; a.first_jet = first_global_jet;
    a.div_magic = D > 1 ? ~0ull / static_cast<unsigned long long>(D) + 1ull : 0ull;
    double run = 0.0;
    for (int m = 0; m <= D; ++m) {                       // cdf in double, rounded once; the last entry is exactly 1
        run += static_cast<double>(mult_probs[m]);
        a.cdf[m] = m == D ? 1.0f : static_cast<float>(run / tot);
    }
    return launch_make_source(a, x0, reinterpret_cast<long long*>(k0), reinterpret_cast<long long*>(mask), n_out,
                              static_cast<cudaStream_t>(stream));
}

int mmf_bridge_sample(const float* x0, const float* x1, const int64_t* k0, const int64_t* k1, const float* t, float sigma, float beta,
                      int32_t V, const float* z, const float* u, uint64_t seed, uint64_t first_global_jet, int32_t B, int32_t D,
                      float* xt, int64_t* kt, int32_t device, void* stream) {
    MMF_REQUIRE(x0 && x1 && k0 && k1 && t && xt && kt, "null argument");
    MMF_REQUIRE(B >= 0 && D >= 1, "bad batch shape");
    MMF_CUDA_OK(cudaSetDevice(device));
    int* flag = nullptr;
    MMF_TRY(step_flag(device, &flag));
    return launch_bridge_sample(x0, x1, reinterpret_cast<const long long*>(k0), reinterpret_cast<const long long*>(k1), t, sigma, beta, V, z, u,
                                seed, first_global_jet * static_cast<uint64_t>(D), B, D, xt, reinterpret_cast<long long*>(kt), flag,
                                static_cast<cudaStream_t>(stream));
}

int mmf_multitask_loss(const float* vt, const float* logits, const float* x0, const float* x1, const int64_t* k1, const int64_t* mask,
                       const float* t, int32_t B, int32_t D, int32_t V, int32_t mode, int32_t n_embd, const float* w_fc, const float* b_fc,
                       const float* w_proj, const float* b_proj, float* per_jet /* 2 B */, float* out5, int32_t device, void* stream) {
    MMF_REQUIRE(vt && logits && x0 && x1 && k1 && mask && t && per_jet && out5, "null argument");
    MMF_REQUIRE(B >= 0 && D >= 1 && (mode == 0 || mode == 1), "bad arguments");
    MMF_CUDA_OK(cudaSetDevice(device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_TRY(launch_multitask_loss(vt, logits, x0, x1, reinterpret_cast<const long long*>(k1), reinterpret_cast<const long long*>(mask), B, D, V,
                                  per_jet, per_jet + B, s));
    return launch_loss_combine(t, per_jet, per_jet + B, w_fc, b_fc, w_proj, b_proj, n_embd, mode, B, out5, s);
}

int mmf_ema_update(float* ema, const float* p, double decay, int64_t n, int32_t device, void* stream) {
    MMF_REQUIRE(n >= 0 && decay >= 0.0 && decay <= 1.0, "bad arguments");
    if (n == 0) return 0;
    MMF_REQUIRE(ema && p, "null argument");
    MMF_CUDA_OK(cudaSetDevice(device));
    return launch_ema_update(ema, p, decay, n, static_cast<cudaStream_t>(stream));
}

int64_t mmf_sample_record_bytes(int32_t D) { return D >= 1 ? sample_record_bytes(D) : 0; }

int mmf_pack_sample(const float* x, const int64_t* k, const int64_t* mask, const float* mean, const float* std_, int64_t B,
                    int32_t D, uint8_t* records, int32_t device, void* stream) {
    MMF_REQUIRE(B >= 0 && D >= 1, "bad sample shape");
    if (B == 0) return 0;
    MMF_REQUIRE(x && mask && records, "null argument");
    MMF_CUDA_OK(cudaSetDevice(device));
    return launch_sample_pack(x, reinterpret_cast<const long long*>(k), reinterpret_cast<const long long*>(mask), mean, std_, B, D,
                              records, static_cast<cudaStream_t>(stream));
}

int mmf_unpack_sample(const uint8_t* records, int64_t B, int32_t D, float* x, int64_t* k, int64_t* mask, int32_t device, void* stream) {
    MMF_REQUIRE(B >= 0 && D >= 1, "bad sample shape");
    if (B == 0) return 0;
    MMF_REQUIRE(records && x, "null argument");
    MMF_CUDA_OK(cudaSetDevice(device));
    return launch_sample_unpack(records, B, D, x, reinterpret_cast<long long*>(k), reinterpret_cast<long long*>(mask),
                                static_cast<cudaStream_t>(stream));
}

int mmf_generate(MmfModel* m, const float* x0, const int64_t* k0, const int64_t* mask, int32_t B, int32_t D,
                 const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, const float* u,
                 const uint8_t* forced_k, float* x_out, int64_t* k_out, float* rates_out, void* stream) {
    MMF_REQUIRE(m && x0 && mask && t_grid && x_out, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_CUDA_OK(cudaSetDevice(m->device));
    std::vector<int64_t> hmask;
    MMF_TRY(fetch_mask(mask, B, D, &hmask, s));
    MMF_TRY(generate_device(m, x0, k0, hmask.data(), B, D, t_grid, N, dt, opts, u, forced_k, x_out, k_out, rates_out, s));
    return check_device_flags(m, s);
}

static int prefix_mask_from_counts(const int32_t* n_per_jet, int B, int D, std::vector<int64_t>* hmask) {
    hmask->assign(static_cast<size_t>(B) * D, 0);
    for (int b = 0; b < B; ++b) {
        MMF_REQUIRE(n_per_jet[b] >= 0 && n_per_jet[b] <= D, "n_per_jet out of [0, D]");
        std::fill(hmask->begin() + static_cast<size_t>(b) * D, hmask->begin() + static_cast<size_t>(b) * D + n_per_jet[b], 1);
    }
    return 0;
}

int mmf_generate_n(MmfModel* m, const float* x0, const int64_t* k0, const int32_t* n_per_jet, int32_t B, int32_t D,
                   const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, const float* u,
                   const uint8_t* forced_k, float* x_out, int64_t* k_out, float* rates_out, void* stream) {
    MMF_REQUIRE(m && x0 && n_per_jet && t_grid && x_out, "null argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    MMF_CUDA_OK(cudaSetDevice(m->device));
    std::vector<int64_t> hmask;
    MMF_TRY(prefix_mask_from_counts(n_per_jet, B, D, &hmask));
    return generate_device(m, x0, k0, hmask.data(), B, D, t_grid, N, dt, opts, u, forced_k, x_out, k_out, rates_out, s);
}

int mmf_model_status(MmfModel* m, void* stream) {
    MMF_REQUIRE(m != nullptr, "null model");
    MMF_CUDA_OK(cudaSetDevice(m->device));
    return check_device_flags(m, static_cast<cudaStream_t>(stream));
}

int mmf_generate_host(MmfModel* m, const float* x0, const int64_t* k0, const int64_t* mask, int32_t B, int32_t D,
                      const float* t_grid, int32_t N, float dt, const MmfStepOptions* opts, float* x_out, int64_t* k_out,
                      void* stream) {
    MMF_REQUIRE(m && x0 && mask && t_grid && x_out, "null argument");
    MMF_CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t slots = static_cast<size_t>(B) * D;
    const size_t need = slots * (3 * 4 + 8);
    Workspace& w = m->ws;
    if (w.stage_bytes < need) {
        if (w.stage) { MMF_CUDA_OK(cudaDeviceSynchronize()); MMF_CUDA_OK(cudaFree(w.stage)); w.stage = nullptr; }
        MMF_CUDA_OK(cudaMalloc(&w.stage, need));
        w.stage_bytes = need;
    }
    float* dx = reinterpret_cast<float*>(w.stage);
    int64_t* dk = reinterpret_cast<int64_t*>(w.stage + slots * 3 * 4);
    MMF_CUDA_OK(cudaMemcpyAsync(dx, x0, slots * 3 * 4, cudaMemcpyHostToDevice, s));
    if (k0) MMF_CUDA_OK(cudaMemcpyAsync(dk, k0, slots * 8, cudaMemcpyHostToDevice, s));
    MMF_TRY(generate_device(m, dx, k0 ? dk : nullptr, mask, B, D, t_grid, N, dt, opts, nullptr, nullptr, dx, k0 ? dk : nullptr, nullptr, s));
    MMF_CUDA_OK(cudaMemcpyAsync(x_out, dx, slots * 3 * 4, cudaMemcpyDeviceToHost, s));
    if (k_out) MMF_CUDA_OK(cudaMemcpyAsync(k_out, dk, slots * 8, cudaMemcpyDeviceToHost, s));
    return check_device_flags(m, s);
}

// ------------------------------------------------------------------------------- diagnostics
int mmf_dbg_ring_plan(const int32_t* tile_kb, int32_t n, int32_t* dst_kb, int32_t* dep) {
    MMF_REQUIRE(tile_kb && dst_kb && dep && n > 0, "dbg_ring_plan: null argument");
    std::vector<int> kb(tile_kb, tile_kb + n), d, q;
    MMF_REQUIRE(plan_weight_ring(kb, &d, &q), "dbg_ring_plan: no valid plan for these tile sizes");
    for (int i = 0; i < n; ++i) { dst_kb[i] = d[i]; dep[i] = q[i]; }
    return 0;
}

int mmf_dbg_gemm(const void* A, const void* W, const float* bias, int32_t M, int32_t N, int32_t K, int32_t mode,
                 int32_t act, void* out, int32_t device, void* stream) {
    MMF_REQUIRE(M % 128 == 0 && N % 128 == 0 && K % 64 == 0, "dbg_gemm: M,N multiples of 128, K multiple of 64");
    MMF_CUDA_OK(cudaSetDevice(device));
    CUtensorMap ta, tb, to;
    MMF_TRY(make_tmap_2d(&ta, A, 2, M, K, K, 64, 128));
    MMF_TRY(make_tmap_2d(&tb, W, 2, N, K, K, 64, 128));
    if (mode == 0) MMF_TRY(make_tmap_2d(&to, out, 2, M, N, N, 64, 128));
    else MMF_TRY(make_tmap_2d(&to, out, 4, M, N, N, 32, 128));
    GemmArgs a{};
    a.kblocks = K / 64; a.w_rows_per_group = N; a.bias = bias; a.act = act;
    return launch_gemm(mode == 0 ? EPI_STORE_BF16 : EPI_STORE_F32, 128, ta, tb, to, to, a, M / 128, N / 128, 1,
                       static_cast<cudaStream_t>(stream));
}

int mmf_dbg_gemm_resln(const void* A, const void* W, const float* bias, const float* temb, const float* ln_g,
                       const float* ln_b, int32_t M, int32_t C, int32_t K, float* residual, void* act_out, int32_t device,
                       void* stream) {
    MMF_REQUIRE(M % 128 == 0 && (C == 128 || C == 256) && K % 64 == 0, "dbg_gemm_resln: bad shape");
    MMF_CUDA_OK(cudaSetDevice(device));
    CUtensorMap ta, tb, to, tr;
    MMF_TRY(make_tmap_2d(&ta, A, 2, M, K, K, 64, 128));
    MMF_TRY(make_tmap_2d(&tb, W, 2, C, K, K, 64, C));
    MMF_TRY(make_tmap_2d(&to, act_out, 2, M, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&tr, residual, 4, M, C, C, 32, 128));
    GemmArgs a{};
    a.kblocks = K / 64; a.w_rows_per_group = C; a.bias = bias; a.temb = temb; a.temb_ld = C; a.ln_g = ln_g; a.ln_b = ln_b;
    return launch_gemm(EPI_RESLN, C, ta, tb, to, tr, a, M / 128, 1, 1, static_cast<cudaStream_t>(stream));
}

int mmf_dbg_gemm_qkv(const void* A, const void* W, const float* bias, const float* q_g, const float* q_b, const float* k_g,
                     const float* k_b, int32_t M, int32_t C, int32_t hs, void* q_out, void* k_out, void* vT_out,
                     int32_t device, void* stream) {
    MMF_REQUIRE(M % 128 == 0 && (C == 128 || C == 256) && (hs == 32 || hs == 64), "dbg_gemm_qkv: bad shape");
    MMF_CUDA_OK(cudaSetDevice(device));
    CUtensorMap ta, tb, tq, tk;
    MMF_TRY(make_tmap_2d(&ta, A, 2, M, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&tb, W, 2, 3 * C, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&tq, q_out, 2, M, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&tk, k_out, 2, M, C, C, 64, 128));
    GemmArgs a{};
    a.kblocks = C / 64; a.w_rows_per_group = 3 * C; a.bias = bias; a.sect_width = C; a.hs = hs;
    a.q_g = q_g; a.q_b = q_b; a.k_g = k_g; a.k_b = k_b; a.vt = static_cast<bf16*>(vT_out); a.vt_ld = M;
    return launch_gemm(EPI_QKV, 128, ta, tb, tq, tk, a, M / 128, 3 * C / 128, 1, static_cast<cudaStream_t>(stream));
}

int mmf_dbg_attention(const void* q, const void* k, const void* vT, const int32_t* jet_n, int32_t n_jets, int32_t M,
                      int32_t C, int32_t hs, void* out, int32_t device, void* stream) {
    MMF_REQUIRE(C % 64 == 0 && M % 8 == 0, "dbg_attention: C multiple of 64, M multiple of 8");
    MMF_CUDA_OK(cudaSetDevice(device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    std::vector<int> start(n_jets), n(n_jets), seg_beg, seg_end;
    int rows = 0;
    for (int j = 0; j < n_jets; ++j) {
        MMF_REQUIRE(jet_n[j] >= 0 && jet_n[j] <= kMaxKeys - 8, "dbg_attention: jet too long");
        start[j] = rows; n[j] = jet_n[j]; rows += jet_n[j];
        for (int i = 0; i < jet_n[j]; ++i) { seg_beg.push_back(start[j]); seg_end.push_back(start[j] + jet_n[j]); }
    }
    MMF_REQUIRE(rows <= M, "dbg_attention: jets exceed M rows");
    std::vector<AttnItem> items;
    plan_attention_items(start, n, &items);
    if (items.empty()) return 0;
    int *d_beg = nullptr, *d_end = nullptr;
    AttnItem* d_items = nullptr;
    MMF_CUDA_OK(cudaMalloc(&d_beg, std::max(rows, 1) * 4));
    MMF_CUDA_OK(cudaMalloc(&d_end, std::max(rows, 1) * 4));
    MMF_CUDA_OK(cudaMalloc(&d_items, items.size() * sizeof(AttnItem)));
    MMF_CUDA_OK(cudaMemcpyAsync(d_beg, seg_beg.data(), rows * 4, cudaMemcpyHostToDevice, s));
    MMF_CUDA_OK(cudaMemcpyAsync(d_end, seg_end.data(), rows * 4, cudaMemcpyHostToDevice, s));
    MMF_CUDA_OK(cudaMemcpyAsync(d_items, items.data(), items.size() * sizeof(AttnItem), cudaMemcpyHostToDevice, s));
    CUtensorMap tq, tk, tv;
    MMF_TRY(make_tmap_2d(&tq, q, 2, M, C, C, 64, 128));
    MMF_TRY(make_tmap_2d(&tk, k, 2, M, C, C, 64, 32));
    MMF_TRY(make_tmap_2d(&tv, vT, 2, C, M, M, 64, 64));
    AttnArgs a{};
    a.items = d_items; a.seg_beg = d_beg; a.seg_end = d_end; a.out = static_cast<bf16*>(out); a.ld_out = C; a.hs = hs;
    a.scale_log2e = 1.4426950408889634f / std::sqrt(static_cast<float>(hs));
    int rc = launch_attention(tq, tk, tv, a, static_cast<int>(items.size()), C / 64, s);
    cudaStreamSynchronize(s);
    cudaFree(d_beg); cudaFree(d_end); cudaFree(d_items);
    return rc;
}

}  // extern "C"
