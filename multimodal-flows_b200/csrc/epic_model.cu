// Host side of the EPiC path: checkpoint folding (weight-norm, wxe and time columns), the weight stream in
// consumption order, tile planning (whole jets per 128-row tile; jets above 128 particles are split over a
// 2-CTA cluster) and the launches.  Reference: networks/EPiC.py, model/CFM.py:133-154.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "mmf_epic.h"
#include "mmf_simt.h"

namespace mmf {

struct EpicModel {
    DeviceArena arena;
    EpicParams p{};
    EpicTimeFold fold{};
    // workspace, grown on demand
    int tile_cap = 0, tb_cap = 0;
    uint8_t* ws = nullptr;
    EpicTileMeta* d_meta = nullptr;
    float *d_xs0 = nullptr, *d_skip = nullptr, *d_tbias = nullptr, *d_temb = nullptr;
    int* d_row_slot = nullptr;
    int64_t launches = 0;
    PinnedStage stage;                                   // per-call tables on their way to the device
    ~EpicModel() {
        stage.release();
        arena.release();
        if (ws) cudaFree(ws);
    }
};

namespace {

// W = g * v / ||v||_row   (old-style torch.nn.utils.weight_norm, dim=0; reference EPiC.py:4,97-106,145-148)
std::vector<float> wn_weight(WeightMap& wm, const std::string& name, int n_out, int n_in) {
    std::vector<float> v = wm.get(name + ".weight_v", n_out, n_in), g = wm.get(name + ".weight_g", n_out, 1);
    for (int o = 0; o < n_out; ++o) {
        double q = 0;
        for (int i = 0; i < n_in; ++i) q += static_cast<double>(v[static_cast<size_t>(o) * n_in + i]) * v[static_cast<size_t>(o) * n_in + i];
        const float norm = static_cast<float>(std::sqrt(q));
        for (int i = 0; i < n_in; ++i) {
            float& x = v[static_cast<size_t>(o) * n_in + i];
            x = g[o] * x / norm;
        }
    }
    return v;
}

// columns [c0, c0+nc) of a row-major [n_out][n_in] matrix
std::vector<float> cols(const std::vector<float>& w, int n_out, int n_in, int c0, int nc) {
    std::vector<float> out(static_cast<size_t>(n_out) * nc);
    for (int o = 0; o < n_out; ++o)
        for (int c = 0; c < nc; ++c) out[static_cast<size_t>(o) * nc + c] = w[static_cast<size_t>(o) * n_in + c0 + c];
    return out;
}
std::vector<float> transpose(const std::vector<float>& w, int n_out, int n_in) {
    std::vector<float> out(w.size());
    for (int o = 0; o < n_out; ++o)
        for (int i = 0; i < n_in; ++i) out[static_cast<size_t>(i) * n_out + o] = w[static_cast<size_t>(o) * n_in + i];
    return out;
}

// append a [256 out][256 in] matrix to the weight stream: for kb, for nh: one 128x64 bf16 tile in SWIZZLE_128B order
void stream_matrix(std::vector<uint16_t>& stream, const std::vector<float>& w /*[256][256]*/) {
    for (int kb = 0; kb < 4; ++kb)
        for (int nh = 0; nh < 2; ++nh) {
            const size_t base = stream.size();
            stream.resize(base + 128 * 64);
            for (int rr = 0; rr < 128; ++rr)
                for (int e = 0; e < 64; ++e) {
                    const size_t off = static_cast<size_t>(rr) * 64 + (((e >> 3) ^ (rr & 7)) << 3) + (e & 7);
                    stream[base + off] = f32_to_bf16_bits(w[static_cast<size_t>(nh * 128 + rr) * 256 + kb * 64 + e]);
                }
        }
}

size_t put_u16(DeviceArena& ar, const std::vector<uint16_t>& v) {
    const size_t off = ar.reserve(v.size() * 2);
    memcpy(ar.staging.data() + off, v.data(), v.size() * 2);
    return off;
}

struct EpicPlan {
    std::vector<EpicTileMeta> meta;
    std::vector<int> row_slot;
    int n_plain = 0, n_pair_tiles = 0;
};

int plan_tiles(const int64_t* mask, int B, int D, bool per_jet_time, EpicPlan* p) {
    std::vector<int> n(B, 0);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) n[b] += mask[static_cast<size_t>(b) * D + d] != 0;
    std::vector<int> order(B);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n[a] > n[b]; });
    struct Bin { int rows = 0; std::vector<int> jets; };
    std::vector<Bin> bins;
    std::vector<int> big;
    for (int b : order) {
        if (n[b] == 0) continue;                         // empty jets produce nothing (EPiC.py:70 divides by zero there)
        if (n[b] > 128) { big.push_back(b); continue; }
        int best = -1;
        for (size_t i = 0; i < bins.size(); ++i)         // best fit: the fullest bin that still takes the jet
            if (bins[i].rows + n[b] <= 128 && static_cast<int>(bins[i].jets.size()) < kEpicMaxJets &&
                (best < 0 || bins[i].rows > bins[best].rows)) best = static_cast<int>(i);
        if (best < 0) { bins.emplace_back(); best = static_cast<int>(bins.size()) - 1; }
        bins[best].rows += n[b];
        bins[best].jets.push_back(b);
    }
    p->meta.clear();
    p->row_slot.clear();
    auto real_slots = [&](int b) {
        std::vector<int> s;
        for (int d = 0; d < D; ++d)
            if (mask[static_cast<size_t>(b) * D + d] != 0) s.push_back(b * D + d);
        return s;
    };
    for (const Bin& bin : bins) {
        EpicTileMeta m{};
        std::vector<int> slots;
        for (size_t j = 0; j < bin.jets.size(); ++j) {
            const int b = bin.jets[j];
            m.jet_begin[j] = static_cast<int>(slots.size());
            m.jet_ntot[j] = n[b];
            m.jet_tb[j] = per_jet_time ? b : 0;
            const std::vector<int> s = real_slots(b);
            slots.insert(slots.end(), s.begin(), s.end());
        }
        m.njets = static_cast<int>(bin.jets.size());
        for (int j = m.njets; j <= kEpicMaxJets; ++j) m.jet_begin[j] = static_cast<int>(slots.size());
        m.nrows = static_cast<int>(slots.size());
        slots.resize(128, -1);
        p->meta.push_back(m);
        p->row_slot.insert(p->row_slot.end(), slots.begin(), slots.end());
    }
    p->n_plain = static_cast<int>(bins.size());
    for (int b : big) {
        const std::vector<int> s = real_slots(b);
        const int h0 = (n[b] + 1) / 2;
        for (int half = 0; half < 2; ++half) {
            EpicTileMeta m{};
            std::vector<int> slots(half == 0 ? s.begin() : s.begin() + h0, half == 0 ? s.begin() + h0 : s.end());
            m.nrows = static_cast<int>(slots.size());
            m.njets = 1;
            m.pair = 1;
            m.jet_begin[0] = 0;
            for (int j = 1; j <= kEpicMaxJets; ++j) m.jet_begin[j] = m.nrows;
            m.jet_ntot[0] = n[b];
            m.jet_tb[0] = per_jet_time ? b : 0;
            slots.resize(128, -1);
            p->meta.push_back(m);
            p->row_slot.insert(p->row_slot.end(), slots.begin(), slots.end());
        }
    }
    p->n_pair_tiles = 2 * static_cast<int>(big.size());
    return 0;
}

int ensure_ws(EpicModel* m, int tiles, int tb) {
    if (tiles <= m->tile_cap && tb <= m->tb_cap) return 0;
    // 25 % headroom on growth: batches of a run differ by a few tiles, and a growth is a device-wide synchronise + free + allocate
    const int tc = tiles > m->tile_cap ? std::max(tiles + tiles / 4, 1) : m->tile_cap, bc = std::max(m->tb_cap, std::max(tb, 1));
    if (m->ws) { MMF_CUDA_OK(cudaDeviceSynchronize()); MMF_CUDA_OK(cudaFree(m->ws)); m->ws = nullptr; }
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    const size_t rows = static_cast<size_t>(tc) * 128;
    const size_t o_meta = take(static_cast<size_t>(tc) * sizeof(EpicTileMeta)), o_xs = take(rows * 3 * 4), o_slot = take(rows * 4),
                 o_skip = take(rows * 256 * 4), o_tb = take(static_cast<size_t>(bc) * kEpicTbLd * 4), o_temb = take(static_cast<size_t>(bc) * 256 * 4);
    MMF_CUDA_OK(cudaMalloc(&m->ws, off));
    m->tile_cap = tc; m->tb_cap = bc;
    m->d_meta = reinterpret_cast<EpicTileMeta*>(m->ws + o_meta);
    m->d_xs0 = reinterpret_cast<float*>(m->ws + o_xs);
    m->d_row_slot = reinterpret_cast<int*>(m->ws + o_slot);
    m->d_skip = reinterpret_cast<float*>(m->ws + o_skip);
    m->d_tbias = reinterpret_cast<float*>(m->ws + o_tb);
    m->d_temb = reinterpret_cast<float*>(m->ws + o_temb);
    return 0;
}

// shared by the forward API (per-jet times) and the sampler (one time per step)
int run(EpicModel* m, const float* x, const int64_t* mask_host, int B, int D, const float* times, int n_times,
        bool per_jet_time, int nsteps, float dt, float* x_out, float* vt_out, cudaStream_t s) {
    EpicPlan plan;
    MMF_TRY_RC(plan_tiles(mask_host, B, D, per_jet_time, &plan));
    const int tiles = static_cast<int>(plan.meta.size());
    MMF_TRY_RC(ensure_ws(m, tiles, n_times));
    float* out = vt_out ? vt_out : x_out;
    if (tiles == 0) {
        MMF_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(B) * D * 3 * 4, s));
        return 0;
    }
    std::vector<float> temb(static_cast<size_t>(n_times) * 256);
    for (int i = 0; i < n_times; ++i) sincos_row(times[i], 256, &temb[static_cast<size_t>(i) * 256]);
    MMF_TRY_RC(m->stage.begin(temb.size() * 4 + plan.meta.size() * sizeof(EpicTileMeta) + plan.row_slot.size() * 4 + 64));
    MMF_TRY_RC(m->stage.push(m->d_temb, temb.data(), temb.size() * 4, s));
    MMF_TRY_RC(m->stage.push(m->d_meta, plan.meta.data(), plan.meta.size() * sizeof(EpicTileMeta), s));
    MMF_TRY_RC(m->stage.push(m->d_row_slot, plan.row_slot.data(), plan.row_slot.size() * 4, s));
    MMF_TRY_RC(m->stage.end(s));                     // pinned staging: no host synchronisation
    MMF_TRY_RC(launch_epic_time_bias(m->fold, m->d_temb, n_times, m->d_tbias, s));
    MMF_TRY_RC(launch_pack(x, nullptr, m->d_row_slot, tiles * 128, 0, m->d_xs0, nullptr, nullptr, s));
    MMF_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(B) * D * 3 * 4, s));     // after the pack: x_out may alias x0
    m->launches += 2;
    EpicLaunch a{};
    a.p = m->p; a.meta = m->d_meta; a.xs0 = m->d_xs0; a.row_slot = m->d_row_slot; a.loc_skip = m->d_skip; a.tbias = m->d_tbias;
    a.per_jet_time = per_jet_time ? 1 : 0; a.nsteps = nsteps; a.dt = dt; a.x_out = vt_out ? nullptr : x_out; a.vt_out = vt_out;
    const char* trace_path = getenv("MMF_TRACE");
    unsigned long long* d_trace = nullptr;
    if (trace_path) {
        MMF_CUDA_OK(cudaMalloc(&d_trace, 128 * 8));
        MMF_CUDA_OK(cudaMemsetAsync(d_trace, 0, 128 * 8, s));
        a.trace = d_trace;
    }
    if (plan.n_plain) {
        a.tile0 = 0;
        MMF_TRY_RC(launch_epic_tiles(a, plan.n_plain, 1, s));
        m->launches += 1;
    }
    if (plan.n_pair_tiles) {
        a.tile0 = plan.n_plain;
        MMF_TRY_RC(launch_epic_tiles(a, plan.n_pair_tiles, 2, s));
        m->launches += 1;
    }
    if (d_trace) {                                   // debugging aid: per-phase clock stamps of CTA 0, first two timesteps
        unsigned long long h[128];
        MMF_CUDA_OK(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, s));
        MMF_CUDA_OK(cudaStreamSynchronize(s));
        cudaFree(d_trace);
        if (FILE* f = fopen(trace_path, "w")) {
            for (int st = 0; st < 2; ++st)
                for (int i = 0; i < 64 && h[st * 64 + i]; ++i)
                    fprintf(f, "step %d mark %2d  +%llu cycles (total %llu)\n", st, i, i ? h[st * 64 + i] - h[st * 64 + i - 1] : 0ull,
                            h[st * 64 + i] - h[st * 64]);
            fclose(f);
        }
    }
    return 0;
}

}  // namespace

int epic_create(const MmfModelDesc& d, WeightMap& wm, EpicModel** out) {
    MMF_REQUIRE(d.n_embd == 256 && d.n_embd_glob == 16 && d.n_layer == kEpicLayers && d.dim_continuous == 3,
                "the EPiC kernel is built for n_embd=256, n_embd_glob=16, n_layer=5, dim_continuous=3");
    std::unique_ptr<EpicModel> m(new EpicModel());
    DeviceArena& ar = m->arena;
    const int E = 256, G = 16;

    const std::vector<float> wxe = wm.get("epic.wxe.weight", E, 3), bxe = wm.get("epic.wxe.bias", E);
    const std::vector<float> w1 = wn_weight(wm, "epic.proj.mlp_local.0", E, 2 * E), b1 = wm.get("epic.proj.mlp_local.0.bias", E);
    const std::vector<float> w2 = wn_weight(wm, "epic.proj.mlp_local.2", E, E), b2 = wm.get("epic.proj.mlp_local.2.bias", E);
    const std::vector<float> wg0 = wn_weight(wm, "epic.proj.mlp_global.0", E, 3 * E), bg0 = wm.get("epic.proj.mlp_global.0.bias", E);
    const std::vector<float> wg2p = wn_weight(wm, "epic.proj.mlp_global.2", G, E), bg2p = wm.get("epic.proj.mlp_global.2.bias", G);
    const std::vector<float> wh = wm.get("epic.head.weight", 3, 2 * E + G), bh = wm.get("epic.head.bias", 3);

    // K = 3 fold of wxe into proj.mlp_local.0:  W1[:, 256:] (Wxe x + bxe) = A3 x + W1[:, 256:] bxe
    std::vector<float> a3(static_cast<size_t>(E) * 3), cst(7 * 256), wt;
    for (int o = 0; o < E; ++o) {
        double c = b1[o];
        double acc[3] = {0, 0, 0};
        for (int i = 0; i < E; ++i) {
            const double w = w1[static_cast<size_t>(o) * 2 * E + E + i];
            c += w * bxe[i];
            for (int q = 0; q < 3; ++q) acc[q] += w * wxe[static_cast<size_t>(i) * 3 + q];
        }
        for (int q = 0; q < 3; ++q) a3[static_cast<size_t>(o) * 3 + q] = static_cast<float>(acc[q]);
        cst[o] = static_cast<float>(c);
    }
    append(wt, transpose(cols(w1, E, 2 * E, 0, E), E, E));                  // m = 0: time columns of proj.mlp_local.0
    append(wt, transpose(cols(wg0, E, 3 * E, 2 * E, E), E, E));             // m = 1: time columns of proj.mlp_global.0
    for (int o = 0; o < E; ++o) cst[256 + o] = bg0[o];

    std::vector<uint16_t> stream;
    stream_matrix(stream, w2);
    size_t o_wg1t[kEpicLayers], o_bg1[kEpicLayers], o_wg2[kEpicLayers], o_bg2[kEpicLayers], o_wl1g[kEpicLayers], o_bl2[kEpicLayers];
    for (int l = 0; l < kEpicLayers; ++l) {
        const std::string q = "epic.layers." + std::to_string(l);
        const std::vector<float> g1 = wn_weight(wm, q + ".fc_glob1", E, 2 * E + G), g2 = wn_weight(wm, q + ".fc_glob2", G, E),
                                 l1 = wn_weight(wm, q + ".fc_loc1", E, 2 * E + G), l2 = wn_weight(wm, q + ".fc_loc2", E, E);
        stream_matrix(stream, cols(l1, E, 2 * E + G, E, E));
        stream_matrix(stream, l2);
        append(wt, transpose(cols(l1, E, 2 * E + G, 0, E), E, E));          // m = 2 + l: time columns of fc_loc1
        const std::vector<float> bl1 = wm.get(q + ".fc_loc1.bias", E);
        for (int o = 0; o < E; ++o) cst[(2 + l) * 256 + o] = bl1[o];
        o_wg1t[l] = ar.put_bf16(transpose(g1, E, 2 * E + G));
        o_bg1[l] = ar.put_f32(wm.get(q + ".fc_glob1.bias", E));
        o_wg2[l] = ar.put_f32(g2);
        o_bg2[l] = ar.put_f32(wm.get(q + ".fc_glob2.bias", G));
        o_wl1g[l] = ar.put_f32(cols(l1, E, 2 * E + G, 2 * E, G));
        o_bl2[l] = ar.put_f32(wm.get(q + ".fc_loc2.bias", E));
    }
    if (!wm.missing.empty()) { set_last_error(wm.missing); return 2; }
    MMF_REQUIRE(stream.size() == static_cast<size_t>(kEpicTilesPerStep) * 128 * 64, "epic: weight stream size");

    const size_t o_stream = put_u16(ar, stream), o_a3 = ar.put_f32(a3), o_b2 = ar.put_f32(b2),
                 o_wg0t = ar.put_bf16(transpose(cols(wg0, E, 3 * E, 0, 2 * E), E, 2 * E)), o_wg2p = ar.put_f32(wg2p),
                 o_bg2p = ar.put_f32(bg2p), o_whl = ar.put_f32(cols(wh, 3, 2 * E + G, E, E)),
                 o_whg = ar.put_f32(cols(wh, 3, 2 * E + G, 2 * E, G)), o_wt = ar.put_f32(wt), o_cst = ar.put_f32(cst),
                 o_wht = ar.put_f32(transpose(cols(wh, 3, 2 * E + G, 0, E), 3, E)), o_bh = ar.put_f32(bh);
    if (ar.upload() != 0) return 1;
    EpicParams& p = m->p;
    p.wstream = ar.at<uint8_t>(o_stream); p.a3 = ar.at<float>(o_a3); p.b_loc2p = ar.at<float>(o_b2);
    p.wg0t = ar.at<bf16>(o_wg0t); p.wg2p = ar.at<float>(o_wg2p); p.bg2p = ar.at<float>(o_bg2p);
    for (int l = 0; l < kEpicLayers; ++l) {
        p.wg1t[l] = ar.at<bf16>(o_wg1t[l]); p.bg1[l] = ar.at<float>(o_bg1[l]); p.wg2[l] = ar.at<float>(o_wg2[l]);
        p.bg2[l] = ar.at<float>(o_bg2[l]); p.wl1g[l] = ar.at<float>(o_wl1g[l]); p.bl2[l] = ar.at<float>(o_bl2[l]);
    }
    p.wh_loc = ar.at<float>(o_whl); p.wh_glob = ar.at<float>(o_whg);
    m->fold.wt = ar.at<float>(o_wt); m->fold.cst = ar.at<float>(o_cst); m->fold.wht = ar.at<float>(o_wht); m->fold.bh = ar.at<float>(o_bh);
    *out = m.release();
    return 0;
}

void epic_destroy(EpicModel* m) { delete m; }
int64_t epic_launches(const EpicModel* m) { return m ? m->launches : 0; }

int epic_forward(EpicModel* m, const float* x, const int64_t* mask_host, const float* t_host, int B, int D, float* vt_out,
                 cudaStream_t s) {
    return run(m, x, mask_host, B, D, t_host, B, true, 1, 0.f, nullptr, vt_out, s);
}

int epic_generate(EpicModel* m, const float* x0, const int64_t* mask_host, int B, int D, const float* t_grid, int N, float dt,
                  float* x_out, cudaStream_t s) {
    return run(m, x0, mask_host, B, D, t_grid, N, false, N, dt, x_out, nullptr, s);
}

}  // namespace mmf
