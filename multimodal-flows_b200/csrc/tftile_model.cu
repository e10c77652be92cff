// Host side of the persistent transformer tile kernel: builds, once per checkpoint, the per-timestep op table,
// the weight stream in consumption order and the parameter blobs; plans tiles of whole jets; launches.
// The op sequence below MUST mirror the epilogue program in kernels_tftile.cu stage by stage.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "mmf_simt.h"
#include "mmf_tftile.h"

namespace mmf {

struct TfTileModel {
    MmfModelDesc desc{};
    DeviceArena arena;
    std::unique_ptr<TfOpTable> optab, optab_pair;       // host copies, passed by value at every launch (plain / pair tiles)
    std::unique_ptr<TfProdTable> prodtab;
    int n_ops = 0, n_ops_pair = 0, n_prod = 0, n_blobs = 0;
    const uint8_t* d_wstream = nullptr;
    const float* d_params = nullptr;
    std::vector<float> time_expand_w, time_expand_b;     // ParticleFormer, host fp32
    // workspace
    int tile_cap = 0, tb_cap = 0;
    uint8_t* ws = nullptr;
    TfTileMeta* d_meta = nullptr;
    float *d_xs0 = nullptr, *d_skip = nullptr, *d_temb = nullptr, *d_thermo = nullptr;
    int *d_ks0 = nullptr, *d_row_slot = nullptr;
    int64_t launches = 0;
    TfLaunch pending{};                                  // prepared by tftile_prepare, consumed by tftile_launch
    int pending_tiles = 0, pending_pair_tiles = 0;      // plain tiles [0, pending_tiles), pair tiles behind them
    cudaStream_t side = nullptr;                         // pair tiles run beside the plain ones (fork / join with events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int cluster = 2;                                     // CTAs sharing one weight stream (MMF_TILE_CLUSTER = 1 | 2 | 4)
    int softmax_nomax = 0;                               // every score of the checkpoint is bounded: softmax without the row maximum
    PinnedStage stage;                                   // per-call tables on their way to the device
    ~TfTileModel() {
        stage.release();
        arena.release();
        if (ws) cudaFree(ws);
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
    }
};

// Ring placement.  Tiles are consumed in order, so writing a tile over older ones only has to wait for the NEWEST tile
// it overlaps (all older ones were consumed before it).  Every tile takes the size-aligned position whose newest
// overlapped tile is oldest: with uniform sizes this is a plain ring; with mixed 16 / 24 / 32 KB tiles it keeps two big
// tiles in flight instead of serialising them on a misaligned offset.  Two timesteps are simulated (the plan repeats
// every timestep); the second gives the steady-state dependency of every tile.
//   kb[i]   size of tile i in KB            dst_kb[i]  offset in the ring in KB
//   dep[i]  tile g (global index over the launch) may be written once all tiles up to g - dep[i] were consumed
bool plan_weight_ring(const std::vector<int>& kb, std::vector<int>* dst_kb, std::vector<int>* dep_out) {
    struct Res { long idx; int beg, end; };
    std::vector<Res> res;
    const long n = static_cast<long>(kb.size());
    long required = -1;                           // newest tile that must have been consumed so far
    std::vector<int> place(static_cast<size_t>(n), -1);
    dst_kb->assign(static_cast<size_t>(n), 0);
    dep_out->assign(static_cast<size_t>(n), 0);
    for (int pass = 0; pass < 2; ++pass) {
        for (long j = 0; j < n; ++j) {
            const int bytes = kb[j] * 1024;
            if (bytes <= 0 || bytes > kTfRingBytes) return false;
            int pos = place[j];
            if (pos < 0) {                        // first pass decides; the second pass must repeat the same offsets
                const int align = bytes > 16384 ? 32768 : (bytes > 8192 ? 16384 : 8192);   // 24 KB tiles take a 32 KB half like the 32 KB ones
                long best = 0;
                for (int cand = 0; cand + bytes <= kTfRingBytes; cand += align) {
                    long newest = -1;
                    for (const Res& r : res)
                        if (r.beg < cand + bytes && cand < r.end) newest = std::max(newest, r.idx);
                    if (pos < 0 || newest < best) { pos = cand; best = newest; }
                }
                place[j] = pos;
            }
            size_t newest = res.size();
            for (size_t i = 0; i < res.size(); ++i)
                if (res[i].beg < pos + bytes && pos < res[i].end) newest = i;
            if (newest != res.size()) {
                required = std::max(required, res[newest].idx);
                res.erase(res.begin(), res.begin() + newest + 1);
            }
            const long g = pass * n + j;
            res.push_back(Res{g, pos, pos + bytes});
            if (static_cast<int>(res.size()) >= kTfRingBars) return false;
            const long dep = required < 0 ? 255 : g - required;
            if (dep < 1 || dep > 255) return false;
            if (pass == 1) {
                if (dep > kTfRingBars - 4) return false;      // the producers rely on dep staying well below the barrier count
                (*dst_kb)[j] = pos / 1024;
                (*dep_out)[j] = static_cast<int>(dep);
            }
        }
    }
    return true;
}

namespace {

// operand arena and scratch columns of the two tile kinds (TfLay<PAIR> in mmf_tftile.h), as run-time values for the emitter
struct LayH {
    bool pair;
    uint32_t chunk, oA, oQ, oK, oVT, oP, oO, oH0, oH1;
    uint16_t cQkv64, cS, cO64;
    int keys;
};
template <bool PAIR>
LayH make_lay() {
    using L = TfLay<PAIR>;
    return LayH{PAIR, L::kChunk, L::oA, L::oQ, L::oK, L::oVT, L::oP, L::oO, L::oH0, L::oH1,
                static_cast<uint16_t>(L::cQkv64), static_cast<uint16_t>(L::cS), static_cast<uint16_t>(L::cO64), static_cast<int>(L::kKeys)};
}

uint32_t desc_lo(uint32_t arena_off) { return (arena_off >> 4) | (1u << 16); }
uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24); }

struct Builder {
    std::vector<TfOp> ops;
    std::vector<uint2> tiles;                        // producer table (ring placement filled in by plan_ring)
    std::vector<uint16_t> stream;
    std::vector<float> params;

    // rows[i] = pointer to the start of a weight row (fp32, K-contiguous); the op covers input columns [k0, k0 + 64 nkt)
    // and A chunks a_off, a_off + 16 KB, ...
    void ring_op(uint32_t a_off, const std::vector<const float*>& rows, int k0, int nkt, uint16_t dcol, int acc, int wait, int signal) {
        const int n = static_cast<int>(rows.size());
        for (int kt = 0; kt < nkt; ++kt) {
            const size_t base = stream.size();
            tiles.push_back(make_uint2(static_cast<uint32_t>(base / 64) | static_cast<uint32_t>(n / 8) << 24, 0u));
            stream.resize(base + static_cast<size_t>(n) * 64);
            for (int rr = 0; rr < n; ++rr)
                for (int e = 0; e < 64; ++e)
                    stream[base + static_cast<size_t>(rr) * 64 + (((e >> 3) ^ (rr & 7)) << 3) + (e & 7)] = f32_to_bf16_bits(rows[rr][k0 + kt * 64 + e]);
        }
        TfOp o{};
        o.a_lo = desc_lo(a_off); o.b_lo = 0; o.idesc = idesc_bf16(n); o.dcol = dcol; o.nkt = static_cast<uint8_t>(nkt | ((signal >> 2) << 7));
        o.flags = static_cast<uint8_t>((acc ? kTfOpAcc : 0u) | (wait ? kTfOpWait : 0u) | kTfOpRing | (static_cast<uint32_t>(signal & 3) << 4));
        ops.push_back(o);
    }
    // both operands in the arena; B advances by 8 KB per k-tile
    // b_mn: B is MN-major (rows = K, N contiguous inside the 128-byte row) - bit 16 of the instruction descriptor
    void smem_op(uint32_t a_off, uint32_t b_off, int n, uint16_t dcol, int nkt, bool half_k, int acc, int wait, int signal, bool b_mn = false,
                 uint32_t pair_flags = 0) {
        TfOp o{};
        o.a_lo = desc_lo(a_off); o.b_lo = desc_lo(b_off); o.idesc = idesc_bf16(n) | (b_mn ? 1u << 16 : 0u);
        o.dcol = static_cast<uint16_t>(dcol | (pair_flags << kTfPairShift));
        o.nkt = static_cast<uint8_t>(nkt | ((signal >> 2) << 7));
        o.flags = static_cast<uint8_t>((acc ? kTfOpAcc : 0u) | (wait ? kTfOpWait : 0u) | (half_k ? kTfOpHalfK : 0u) | (b_mn ? kTfOpBMn : 0u) |
                                       kTfOpAttn | (static_cast<uint32_t>(signal & 3) << 4));
        ops.push_back(o);
    }
    // pair tiles: a pure hand-off of the attention issuer (no MMA): the K rows of a unit staged ahead of time
    void pair_send_k() { smem_op(0, 0, 16, 0, 0, false, 0, 1, 0, false, kTfPairSendK); }
    bool plan_ring() {
        std::vector<int> kb(tiles.size());
        for (size_t i = 0; i < tiles.size(); ++i) kb[i] = static_cast<int>(tiles[i].x >> 24);
        std::vector<int> dst, dep;
        if (!plan_weight_ring(kb, &dst, &dep)) return false;
        for (size_t i = 0; i < tiles.size(); ++i) tiles[i].y = static_cast<uint32_t>(dst[i]) | static_cast<uint32_t>(dep[i]) << 8;
        return true;
    }
    float* blob(int idx) {
        if (params.size() < static_cast<size_t>(idx + 1) * kTfParamFloats) params.resize(static_cast<size_t>(idx + 1) * kTfParamFloats, 0.f);
        return params.data() + static_cast<size_t>(idx) * kTfParamFloats;
    }
};

struct Mat {                                         // fp32 [rows][cols] copy of a checkpoint matrix
    std::vector<float> w;
    int rows = 0, cols = 0;
    const float* row(int r) const { return w.data() + static_cast<size_t>(r) * cols; }
};
Mat mat(WeightMap& wm, const std::string& name, int rows, int cols) {
    Mat m;
    m.w = wm.get(name, rows, cols);
    m.rows = rows; m.cols = cols;
    return m;
}
std::vector<const float*> rows_of(const Mat& m, int r0, int n) {
    std::vector<const float*> v(n);
    for (int i = 0; i < n; ++i) v[i] = m.row(r0 + i);
    return v;
}
void put(float* dst, const std::vector<float>& src) { std::copy(src.begin(), src.end(), dst); }

float gelu_h(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// The affine part of a LayerNorm that feeds a linear layer moves into that layer (fp re-association only):
//   W (g * xhat + beta) + b = (W diag(g)) xhat + (b + W beta)
// so the kernel's normalise pass writes xhat itself (ln_to_abuf).  `bias` (in / out): the layer's bias; g / beta: [W.cols].
void fold_ln(Mat& W, std::vector<float>& bias, const std::vector<float>& g, const std::vector<float>& beta) {
    for (int o = 0; o < W.rows; ++o) {
        float* w = W.w.data() + static_cast<size_t>(o) * W.cols;
        double acc = 0.0;
        for (int i = 0; i < W.cols; ++i) { acc += static_cast<double>(w[i]) * beta[i]; w[i] *= g[i]; }
        bias[o] += static_cast<float>(acc);
    }
}

// upper bound of |q . k| * log2(e) / sqrt(hs) for per-head LayerNorm outputs q, k (|xhat|_2 <= sqrt(hs)), bf16 rounding included
double score_bound(WeightMap& wm, const std::string& p, int hs) {
    auto side = [&](const std::string& n) {
        const std::vector<float> g = wm.get(p + ".attn." + n + ".weight", hs), b = wm.get(p + ".attn." + n + ".bias", hs, -1, true);
        double gm = 0.0, b2 = 0.0;
        for (int i = 0; i < hs; ++i) { gm = std::max(gm, static_cast<double>(std::fabs(g[i]))); b2 += static_cast<double>(b[i]) * b[i]; }
        return 1.01 * (gm * std::sqrt(static_cast<double>(hs)) + std::sqrt(b2));
    };
    return side("q_layernorm") * side("k_layernorm") * 1.4426950408889634 / std::sqrt(static_cast<double>(hs));
}

// c_attn / c_fc carry ln1 / ln2 (fold_ln); the v bias travels through the attention (the probabilities of a row sum to one:
// P (V + b_v) = P V + b_v) and the projection into the projection bias: b_proj' = b_proj + W_proj b_v.  The kernel adds no v bias.
// q_g / q_b: the affine part of the q LayerNorm times log2(e) / sqrt(hs) - the scores leave the tensor core in the units of
// the softmax exponent.
struct BlockW { Mat attn, proj, fc, p2; std::vector<float> b_attn, b_fc, b_proj, q_g, q_b; };
BlockW load_block(WeightMap& wm, const std::string& p, int C, int I, int hs) {
    BlockW w{mat(wm, p + ".attn.c_attn.weight", 3 * C, C), mat(wm, p + ".attn.c_proj.weight", C, C),
             mat(wm, p + ".ffw.c_fc.weight", I, C), mat(wm, p + ".ffw.c_proj.weight", C, I),
             wm.get(p + ".attn.c_attn.bias", 3 * C, -1, true), wm.get(p + ".ffw.c_fc.bias", I, -1, true),
             wm.get(p + ".attn.c_proj.bias", C, -1, true), wm.get(p + ".attn.q_layernorm.weight", hs),
             wm.get(p + ".attn.q_layernorm.bias", hs, -1, true)};
    fold_ln(w.attn, w.b_attn, wm.get(p + ".ln1.weight", C), wm.get(p + ".ln1.bias", C, -1, true));
    fold_ln(w.fc, w.b_fc, wm.get(p + ".ln2.weight", C), wm.get(p + ".ln2.bias", C, -1, true));
    // q and k rows centred per head: the per-head LayerNorm that follows does not see a constant added to all hs outputs of a
    // head, and with centred rows (and biases) the head's pre-activations sum to zero - the kernel's LayerNorm skips the mean
    // (ln_regs_centred in kernels_tftile.cu)
    for (int part = 0; part < 2; ++part)
        for (int r0 = part * C; r0 < (part + 1) * C; r0 += hs) {
            for (int i = 0; i < C; ++i) {
                double m = 0.0;
                for (int r = r0; r < r0 + hs; ++r) m += w.attn.w[static_cast<size_t>(r) * C + i];
                m /= hs;
                for (int r = r0; r < r0 + hs; ++r) w.attn.w[static_cast<size_t>(r) * C + i] -= static_cast<float>(m);
            }
            double mb = 0.0;
            for (int r = r0; r < r0 + hs; ++r) mb += w.b_attn[r];
            mb /= hs;
            for (int r = r0; r < r0 + hs; ++r) w.b_attn[r] -= static_cast<float>(mb);
        }
    for (int o = 0; o < C; ++o) {
        double acc = 0.0;
        for (int i = 0; i < C; ++i) acc += static_cast<double>(w.proj.row(o)[i]) * w.b_attn[2 * C + i];
        w.b_proj[o] += static_cast<float>(acc);
    }
    const float c = static_cast<float>(1.4426950408889634 / std::sqrt(static_cast<double>(hs)));
    for (int i = 0; i < hs; ++i) { w.q_g[i] *= c; w.q_b[i] *= c; }
    for (float& x : w.p2.w) x *= 0.5f;                 // the kernel's MLP epilogue produces 2 GELU (one multiply less per pair)
    return w;
}

// MLP of one group.  The up-projection runs as two N = 256 halves over the whole scratch (one MMA instruction per K = 16
// step costs ~90 cycles to issue whatever N is, so wide instructions halve the issue time); the epilogue turns each half
// into two 128-column quarters H0 | H1, and the down-projection of a quarter accumulates onto the residual (N = C).
//   MMA order   fc(h0) | out(q0) | fc(h1) out(q1) | out(q2) | out(q3)      ("|" = waits for the next go of the epilogue)
//   signals     fc(h) -> done[0]; out(q1) -> done[1] (H1 may be rewritten); the last out(q3) -> done[0] when final_signal
void emit_mlp(Builder& b, const LayH& L, const BlockW& w, int C, int a_chunk0, uint16_t dcol_out, bool first_wait, bool final_signal) {
    const int kbC = C / 64;
    auto fc = [&](int h, int wait) { b.ring_op(L.oA + a_chunk0 * L.chunk, rows_of(w.fc, h * 256, 256), 0, kbC, 256, 0, wait, 1); };
    auto out = [&](int q, int wait, int signal) { b.ring_op((q & 1) ? L.oH1 : L.oH0, rows_of(w.p2, 0, C), q * 128, 2, dcol_out, 1, wait, signal); };
    fc(0, first_wait);
    out(0, 1, 0);
    fc(1, 1);
    out(1, 0, 2);
    out(2, 1, 0);
    out(3, 1, final_signal ? 1 : 0);
}

}  // namespace

// The per-timestep program of one tile kind: MMA ops, weight stream (consumption order) and parameter blobs.  Both kinds
// consume the SAME weight tiles in the SAME order (the pair program differs only in operand addresses, scratch columns and
// the attention products), so the stream, ring plan and blobs of the plain program serve both.
static int emit_program(const MmfModelDesc& d, WeightMap& wm, const LayH& L, Builder& b, int* n_blobs, double* score_max) {
    const bool pf = d.arch == MMF_ARCH_PARTICLEFORMER;
    const int E = d.n_embd, h = E / 2, I = d.n_inner, V = d.vocab_size;
    const std::string t = "transformer.";
    const int n_stream = pf ? d.n_layer : 0, n_main = pf ? d.n_layer_fused : d.n_layer;
    const uint32_t oA = L.oA, oQ = L.oQ, oK = L.oK, oVT = L.oVT, oO = L.oO, oP = L.oP, kT = L.chunk;
    int blob_idx = 0;

    // ---------------- embedding stage
    {
        float* P = b.blob(blob_idx);
        const std::vector<float> w0 = wm.get(t + "wxe.0.weight", E, 3), b0 = wm.get(t + "wxe.0.bias", E);
        for (int c = 0; c < E; ++c) { P[tfp::EA_W0 + c * 4] = w0[c * 3]; P[tfp::EA_W0 + c * 4 + 1] = w0[c * 3 + 1]; P[tfp::EA_W0 + c * 4 + 2] = w0[c * 3 + 2]; P[tfp::EA_W0 + c * 4 + 3] = b0[c]; }
        put(P + tfp::EA_BXE2, wm.get(t + "wxe.2.bias", h));
        put(P + tfp::EA_LN1X_G, wm.get(t + "ln1_x.weight", h));
        put(P + tfp::EA_LN1X_B, wm.get(t + "ln1_x.bias", h, -1, true));
        ++blob_idx;
        P = b.blob(blob_idx);
        // discrete embedding branch folded into a V x 128 table (reference ParticleTransformers.py:95-96, 190-191)
        const std::vector<float> emb = wm.get(t + "wye.0.weight", V, E), w2 = wm.get(t + "wye.2.weight", h, E), b2 = wm.get(t + "wye.2.bias", h),
                                 g = wm.get(t + "ln1_y.weight", h), bb = wm.get(t + "ln1_y.bias", h, -1, true);
        for (int k = 0; k < V; ++k) {
            std::vector<double> row(h);
            for (int o = 0; o < h; ++o) {
                double acc = b2[o];
                for (int i = 0; i < E; ++i) acc += static_cast<double>(gelu_h(emb[static_cast<size_t>(k) * E + i])) * w2[static_cast<size_t>(o) * E + i];
                row[o] = static_cast<float>(acc);
            }
            double s = 0, q = 0;
            for (int o = 0; o < h; ++o) s += row[o];
            const double mean = s / h;
            for (int o = 0; o < h; ++o) q += (row[o] - mean) * (row[o] - mean);
            const double rstd = 1.0 / std::sqrt(q / h + 1e-5);
            for (int o = 0; o < h; ++o) P[tfp::EB_YTAB + k * 128 + o] = static_cast<float>((row[o] - mean) * rstd) * g[o] + bb[o];
        }
        if (pf) {
            put(P + tfp::EB_LNN_G, wm.get(t + "blocks_x.0.ln1.weight", h)); put(P + tfp::EB_LNN_G + 128, wm.get(t + "blocks_y.0.ln1.weight", h));
            put(P + tfp::EB_LNN_B, wm.get(t + "blocks_x.0.ln1.bias", h, -1, true)); put(P + tfp::EB_LNN_B + 128, wm.get(t + "blocks_y.0.ln1.bias", h, -1, true));
        } else {
            put(P + tfp::EB_LNN_G, wm.get(t + "blocks.0.ln1.weight", E)); put(P + tfp::EB_LNN_B, wm.get(t + "blocks.0.ln1.bias", E, -1, true));
        }
        ++blob_idx;
        const Mat wxe2 = mat(wm, t + "wxe.2.weight", h, E);
        b.ring_op(oA, rows_of(wxe2, 0, 128), 0, 4, 256, 0, 1, 1);
    }

    // ---------------- stream blocks (ParticleFormer): groups x | y, C = 128, head size 32, units = head pairs
    for (int i = 0; i < n_stream; ++i) {
        const std::string px[2] = {t + "blocks_x." + std::to_string(i), t + "blocks_y." + std::to_string(i)};
        BlockW w[2] = {load_block(wm, px[0], h, I, 32), load_block(wm, px[1], h, I, 32)};
        *score_max = std::max(*score_max, std::max(score_bound(wm, px[0], 32), score_bound(wm, px[1], 32)));
        const bool last = i + 1 == n_stream;
        {
            float* P = b.blob(blob_idx);                 // attention stage (prefetched by the previous stage)
            for (int g = 0; g < 2; ++g) {
                float* G = P + g * tfp::SA_GROUP;
                put(G + tfp::SA_BQKV, w[g].b_attn);
                put(G + tfp::SA_QG, w[g].q_g); put(G + tfp::SA_QB, w[g].q_b);
                put(G + tfp::SA_KG, wm.get(px[g] + ".attn.k_layernorm.weight", 32)); put(G + tfp::SA_KB, wm.get(px[g] + ".attn.k_layernorm.bias", 32, -1, true));
                put(G + tfp::SA_BPROJ, w[g].b_proj);
                put(G + tfp::SA_LN2G, wm.get(px[g] + ".ln2.weight", h)); put(G + tfp::SA_LN2B, wm.get(px[g] + ".ln2.bias", h, -1, true));
            }
            P = b.blob(blob_idx + 1);                    // MLP stage
            for (int g = 0; g < 2; ++g) {
                float* G = P + g * tfp::SM_GROUP;
                put(G + tfp::SM_BFC, w[g].b_fc);
                put(G + tfp::SM_BP2, wm.get(px[g] + ".ffw.c_proj.bias", h, -1, true));
            }
            const std::string nx = last ? t + "ln2_x" : t + "blocks_x." + std::to_string(i + 1) + ".ln1";
            const std::string ny = last ? t + "ln2_y" : t + "blocks_y." + std::to_string(i + 1) + ".ln1";
            put(P + tfp::SM_LNN_G, wm.get(nx + ".weight", h)); put(P + tfp::SM_LNN_G + 128, wm.get(ny + ".weight", h));
            put(P + tfp::SM_LNN_B, wm.get(nx + ".bias", h, -1, true)); put(P + tfp::SM_LNN_B + 128, wm.get(ny + ".bias", h, -1, true));
            if (last) {
                put(P + tfp::SM_LN2ND_G, wm.get(t + "blocks_fuse.0.ln1.weight", E)); put(P + tfp::SM_LN2ND_B, wm.get(t + "blocks_fuse.0.ln1.bias", E, -1, true));
            }
        }
        ++blob_idx;                                      // -> this block's MLP blob
        // Scratch columns of a unit: S_h0 [256,384) S_h1 [384,512); O_h0 [256,288) O_h1 [288,320) (over the consumed S_h0);
        // Q|K [384,512), V [320,384).  The QKV product of the NEXT unit is issued right after the last P V of this one: by
        // then both score tiles are consumed and the O columns are not touched, so it runs under the epilogue of this unit.
        auto qkv = [&](int g, int u, int wait) {
            // rows of c_attn: q [0,128) k [128,256) v [256,384); unit u = heads 2u, 2u+1 = columns u*64..u*64+63
            // one N = 192 MMA per k-tile: rows v | q | k -> columns [320,384) [384,448) [448,512)
            std::vector<const float*> r = rows_of(w[g].attn, 256 + u * 64, 64), qq = rows_of(w[g].attn, u * 64, 64), kk = rows_of(w[g].attn, 128 + u * 64, 64);
            r.insert(r.end(), qq.begin(), qq.end());
            r.insert(r.end(), kk.begin(), kk.end());
            b.ring_op(oA + 2 * g * kT, r, 0, 2, 320, 0, wait, 2);
        };
        qkv(0, 0, 1);
        for (int g = 0; g < 2; ++g)
            for (int u = 0; u < 2; ++u) {
                if (!L.pair) {
                    b.smem_op(oQ, oK, 128, 256, 1, true, 0, 1, 0);               // S of head 0 of the pair
                    b.smem_op(oQ + 64, oK + 64, 128, 384, 1, true, 0, 0, 1);     // S of head 1
                    b.smem_op(oQ, oVT, 32, 256, 2, false, 0, 1, 4, true);        // O_h0 = P_h0 V_h0 (keys 0..63, 64..127); V is [key][d]; -> done[3]
                    b.smem_op(oQ, oVT + 64, 32, 288, 2, false, 0, 1, 1, true);   // O_h1: d columns 32..63 of the V rows
                } else {
                    // pair tile: 160 keys, one head at a time on the score columns [320,480); O_h0 [256,288), O_h1 [288,320);
                    // P [rows][160 keys] = two 64-key k-tiles + one 32-key tail, V rows 128 bytes apart.  K / V exchange hand-offs:
                    // K of the block's first unit is sent by its own hand-off op, K of every later unit right after this unit's
                    // last P V hand-off (the epilogue stages it under that product).
                    if (g == 0 && u == 0) b.pair_send_k();
                    b.smem_op(oQ, oK, 160, L.cS, 1, true, 0, 1, 1, false, kTfPairSendVWaitK);          // S of head 0
                    b.smem_op(oQ + 64, oK + 64, 160, L.cS, 1, true, 0, 1, 1, false, kTfPairFreeK);     // S of head 1 (head 0's scores are in registers)
                    b.smem_op(oP, oVT, 32, 256, 2, false, 0, 1, 0, true, kTfPairWaitV);                // O_h0: keys 0..127
                    b.smem_op(oP + 2 * kT, oVT + 128 * 128, 32, 256, 1, true, 1, 0, 4, true);          //       keys 128..159 -> done[3]
                    b.smem_op(oP, oVT + 64, 32, 288, 2, false, 0, 1, 0, true);                         // O_h1
                    b.smem_op(oP + 2 * kT, oVT + 128 * 128 + 64, 32, 288, 1, true, 1, 0, 1, true, kTfPairFreeV);
                    if (!(g == 1 && u == 1)) b.pair_send_k();
                }
                if (!(g == 1 && u == 1)) qkv(u == 1 ? 1 : g, u == 1 ? 0 : 1, 1);    // released as soon as both score tiles are in registers
                b.ring_op(oO, rows_of(w[g].proj, 0, 128), u * 64, 1, static_cast<uint16_t>(g * 128), 1, 1, (g == 1 && u == 1) ? 1 : 3);   // 3: done[2] = oO may be rewritten
            }
        ++blob_idx;
        for (int g = 0; g < 2; ++g) emit_mlp(b, L, w[g], 128, 2 * g, static_cast<uint16_t>(g * 128), g == 0, g == 1);
    }

    // ---------------- main blocks: C = 256, head size 64, units = heads
    for (int j = 0; j < n_main; ++j) {
        const std::string p = t + (pf ? "blocks_fuse." : "blocks.") + std::to_string(j);
        const BlockW w = load_block(wm, p, E, I, 64);
        *score_max = std::max(*score_max, score_bound(wm, p, 64));
        const bool last = j + 1 == n_main;
        {
            float* P = b.blob(blob_idx);                 // attention stage
            put(P + tfp::BA_BQKV, w.b_attn);
            put(P + tfp::BA_QG, w.q_g); put(P + tfp::BA_QB, w.q_b);
            put(P + tfp::BA_KG, wm.get(p + ".attn.k_layernorm.weight", 64)); put(P + tfp::BA_KB, wm.get(p + ".attn.k_layernorm.bias", 64, -1, true));
            put(P + tfp::BA_BPROJ, w.b_proj);
            put(P + tfp::BA_LN2G, wm.get(p + ".ln2.weight", E)); put(P + tfp::BA_LN2B, wm.get(p + ".ln2.bias", E, -1, true));
            P = b.blob(blob_idx + 1);                    // MLP stage
            put(P + tfp::BM_BFC, w.b_fc);
            put(P + tfp::BM_BP2, wm.get(p + ".ffw.c_proj.bias", E, -1, true));
            if (!last) {
                const std::string nx = t + (pf ? "blocks_fuse." : "blocks.") + std::to_string(j + 1) + ".ln1";
                put(P + tfp::BM_LNN_G, wm.get(nx + ".weight", E)); put(P + tfp::BM_LNN_B, wm.get(nx + ".bias", E, -1, true));
            } else if (pf) {
                put(P + tfp::BM_LNN_G, wm.get(t + "ln3_x.weight", h)); put(P + tfp::BM_LNN_G + 128, wm.get(t + "ln3_y.weight", h));
                put(P + tfp::BM_LNN_B, wm.get(t + "ln3_x.bias", h, -1, true)); put(P + tfp::BM_LNN_B + 128, wm.get(t + "ln3_y.bias", h, -1, true));
            } else {
                put(P + tfp::BM_LNN_G, wm.get(t + "ln2.weight", E)); put(P + tfp::BM_LNN_B, wm.get(t + "ln2.bias", E, -1, true));
            }
        }
        ++blob_idx;                                      // -> this block's MLP blob
        // Scratch columns of a unit: Q|K [256,384), V [384,448); S [256,384); O [448,512).  QKV of the next unit is issued
        // right after P V (S is consumed, O is elsewhere) and runs under the O / projection epilogue of this unit.
        auto qkv = [&](int u, int wait) {
            // one N = 192 MMA per k-tile: rows q | k | v -> columns [256,320) [320,384) [384,448)
            std::vector<const float*> r = rows_of(w.attn, u * 64, 64), kk = rows_of(w.attn, 256 + u * 64, 64), vv = rows_of(w.attn, 512 + u * 64, 64);
            r.insert(r.end(), kk.begin(), kk.end());
            r.insert(r.end(), vv.begin(), vv.end());
            b.ring_op(oA, r, 0, 4, L.cQkv64, 0, wait, 2);      // (pair tiles: columns [320,512), O of the unit lives in [256,320))
        };
        qkv(0, 1);
        for (int u = 0; u < 4; ++u) {
            if (!L.pair) {
                b.smem_op(oQ, oK, 128, 256, 1, false, 0, 1, 1);                  // S = Q K^T
                b.smem_op(oQ, oVT, 64, 448, 2, false, 0, 1, 1, true);            // O = P V; V is [key][d] (MN-major B)
            } else {
                if (u == 0) b.pair_send_k();
                b.smem_op(oQ, oK, 160, L.cS, 1, false, 0, 1, 1, false, kTfPairSendVWaitK | kTfPairFreeK);   // S over 160 keys
                b.smem_op(oP, oVT, 64, L.cO64, 2, false, 0, 1, 0, true, kTfPairWaitV);                        // O: keys 0..127
                b.smem_op(oP + 2 * kT, oVT + 128 * 128, 64, L.cO64, 1, true, 1, 0, 1, true, kTfPairFreeV);    //    keys 128..159
                if (u < 3) b.pair_send_k();
            }
            if (u < 3) qkv(u + 1, 1);                                      // released as soon as the score tile is in registers
            b.ring_op(oO, rows_of(w.proj, 0, 256), u * 64, 1, 0, 1, 1, u == 3 ? 1 : 3);      // N = 256; 3: done[2] = oO may be rewritten
        }
        ++blob_idx;
        emit_mlp(b, L, w, 256, 0, 0, true, true);
    }

    // ---------------- heads: Linear(128,512) + GELU on tensor cores, Linear(512, 3 | V) on CUDA cores
    {
        Mat hx = mat(wm, t + "head_x.0.weight", I, h), hy = mat(wm, t + "head_y.0.weight", I, h);
        std::vector<float> bx0 = wm.get(t + "head_x.0.bias", I), by0 = wm.get(t + "head_y.0.bias", I);
        {   // the last LayerNorm (ParticleFormer ln3_x | ln3_y, fused encoder ln2 = columns [0,128) | [128,256)) folds into the heads
            std::vector<float> gx, bx, gy, by;
            if (pf) {
                gx = wm.get(t + "ln3_x.weight", h); bx = wm.get(t + "ln3_x.bias", h, -1, true);
                gy = wm.get(t + "ln3_y.weight", h); by = wm.get(t + "ln3_y.bias", h, -1, true);
            } else {
                const std::vector<float> g2 = wm.get(t + "ln2.weight", E), b2 = wm.get(t + "ln2.bias", E, -1, true);
                gx.assign(g2.begin(), g2.begin() + h); bx.assign(b2.begin(), b2.begin() + h);
                gy.assign(g2.begin() + h, g2.end()); by.assign(b2.begin() + h, b2.end());
            }
            fold_ln(hx, bx0, gx, bx);
            fold_ln(hy, by0, gy, by);
        }
        std::vector<float> wx2 = wm.get(t + "head_x.2.weight", 3, I), wy2 = wm.get(t + "head_y.2.weight", V, I);
        const std::vector<float> bx2 = wm.get(t + "head_x.2.bias", 3), by2 = wm.get(t + "head_y.2.bias", V);
        for (float& x : wx2) x *= 0.5f;                // the head epilogue produces 2 GELU as well
        for (float& x : wy2) x *= 0.5f;
        float* P = b.blob(blob_idx);
        put(P + tfp::HX_BIAS, bx0); put(P + tfp::HX_W2, wx2); put(P + tfp::HX_B2, bx2);
        for (int q = 0; q < 4; ++q) {
            float* Q = b.blob(blob_idx + 1 + q);
            for (int i = 0; i < 128; ++i) Q[tfp::HY_BIAS + i] = by0[q * 128 + i];
            for (int v = 0; v < V; ++v)
                for (int i = 0; i < 128; ++i) Q[tfp::HY_W2 + v * 128 + i] = wy2[static_cast<size_t>(v) * I + q * 128 + i];
            if (q == 0) for (int v = 0; v < V; ++v) Q[tfp::HY_B2 + v] = by2[v];
        }
        for (int hq = 0; hq < 8; ++hq) {
            const Mat& w = hq < 4 ? hx : hy;
            const int chunk0 = hq < 4 ? 0 : 2;
            b.ring_op(oA + chunk0 * kT, rows_of(w, (hq & 3) * 128, 128), 0, 2, static_cast<uint16_t>(256 + (hq & 1) * 128), 0, hq != 1, 1 + (hq & 1));
        }
        blob_idx += 5;
    }
    if (!wm.missing.empty()) { set_last_error(wm.missing); return 2; }
    b.params.resize(static_cast<size_t>(blob_idx) * kTfParamFloats, 0.f);
    *n_blobs = blob_idx;
    return 0;
}

int tftile_create(const MmfModelDesc& d, WeightMap& wm, TfTileModel** out) {
    *out = nullptr;
    const bool pf = d.arch == MMF_ARCH_PARTICLEFORMER;
    const int E = d.n_embd, h = E / 2, I = d.n_inner, V = d.vocab_size;
    if (!(E == 256 && I == 512 && d.n_head == 4 && V == 9 && d.qk_layernorm)) return 0;     // outside the tile kernel's envelope
    const std::string t = "transformer.";
    std::unique_ptr<TfTileModel> m(new TfTileModel());
    m->desc = d;
    Builder b, bp;
    int blob_idx = 0, blob_idx_pair = 0;
    double score_max = 0.0;
    MMF_TRY_RC(emit_program(d, wm, make_lay<false>(), b, &blob_idx, &score_max));
    MMF_TRY_RC(emit_program(d, wm, make_lay<true>(), bp, &blob_idx_pair, &score_max));
    // |score| log2(e) / sqrt(hs) <= 64 for every block: 2^(+-64) and sums of 160 such terms are far inside fp32 / bf16 range, so
    // the softmax needs no row maximum (kernels_tftile.cu softmax_probs).  MMF_TILE_SOFTMAX_MAX=1 keeps the maximum (tests).
    const char* force_max = getenv("MMF_TILE_SOFTMAX_MAX");
    m->softmax_nomax = (score_max <= 64.0 && !(force_max && force_max[0] == '1')) ? 1 : 0;
    if (getenv("MMF_TILE_DEBUG")) fprintf(stderr, "tile kernel: score bound %.2f (log2 units), softmax %s the row maximum\n", score_max, m->softmax_nomax ? "without" : "with");
    MMF_REQUIRE(bp.tiles.size() == b.tiles.size() && bp.stream == b.stream && blob_idx_pair == blob_idx,
                "tile kernel: the pair program must consume the weight stream of the plain program");
    if (pf) {
        m->time_expand_w = wm.get(t + "time_expand.weight", E, h);
        m->time_expand_b = wm.get(t + "time_expand.bias", E);
    }

    DeviceArena& ar = m->arena;
    MMF_REQUIRE(b.ops.size() <= static_cast<size_t>(kTfMaxOps) && bp.ops.size() <= static_cast<size_t>(kTfMaxOps) &&
                    b.tiles.size() <= static_cast<size_t>(kTfMaxOps),
                "tile kernel: op table too long for the kernel parameter space");
    m->optab.reset(new TfOpTable());
    m->optab_pair.reset(new TfOpTable());
    m->prodtab.reset(new TfProdTable());
    memset(m->optab.get(), 0, sizeof(TfOpTable));
    memset(m->optab_pair.get(), 0, sizeof(TfOpTable));
    memset(m->prodtab.get(), 0, sizeof(TfProdTable));
    MMF_REQUIRE(b.plan_ring(), "tile kernel: weight ring plan failed");
    std::copy(b.ops.begin(), b.ops.end(), m->optab->ops);
    std::copy(bp.ops.begin(), bp.ops.end(), m->optab_pair->ops);
    std::copy(b.tiles.begin(), b.tiles.end(), m->prodtab->e);
    m->n_ops = static_cast<int>(b.ops.size());
    m->n_ops_pair = static_cast<int>(bp.ops.size());
    m->n_prod = static_cast<int>(b.tiles.size());
    m->n_blobs = blob_idx;
    const size_t o_stream = ar.reserve(b.stream.size() * 2);
    memcpy(ar.staging.data() + o_stream, b.stream.data(), b.stream.size() * 2);
    const size_t o_params = ar.put_f32(b.params);
    if (ar.upload() != 0) return 1;
    m->d_wstream = ar.at<uint8_t>(o_stream);
    m->d_params = ar.at<float>(o_params);
    if (const char* c = getenv("MMF_TILE_CLUSTER")) {
        const int v = atoi(c);
        if (v == 1 || v == 2 || v == 4) m->cluster = v;
    }
    *out = m.release();
    return 0;
}

void tftile_destroy(TfTileModel* m) { delete m; }
int64_t tftile_launches(const TfTileModel* m) { return m ? m->launches : 0; }

namespace {

struct TilePlan {
    std::vector<TfTileMeta> meta;        // plain tiles first (padded to a whole number of clusters), then pair tiles
    std::vector<int> row_slot;
    int n_plain = 0, n_pair = 0;
};

constexpr int kPairRows = static_cast<int>(TfLay<true>::kRows);       // rows a CTA of a pair holds at most

void plan_tiles(const int64_t* mask, int B, int D, bool per_jet_time, int cluster, TilePlan* p, std::vector<unsigned char>* handled) {
    std::vector<int> n(B, 0);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) n[b] += mask[static_cast<size_t>(b) * D + d] != 0;
    std::vector<int> order(B);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n[a] > n[b]; });
    struct Bin { int rows = 0; std::vector<int> jets; };
    std::vector<Bin> bins;
    std::vector<int> big;                                             // jets of 129 ... 160 particles: one CTA pair each
    // best-fit decreasing with an index of open bins by free space (128 buckets): O(B * 128)
    std::vector<std::vector<int>> by_free(129);
    handled->assign(B, 0);
    for (int b : order) {
        if (n[b] == 0) { (*handled)[b] = 1; continue; }            // nothing to generate
        if (n[b] > 2 * kPairRows) continue;                          // (not reachable with max_num_particles <= 152)
        (*handled)[b] = 1;
        if (n[b] > 128) { big.push_back(b); continue; }
        int pick = -1;
        for (int f = n[b]; f <= 128 && pick < 0; ++f)
            if (!by_free[f].empty()) { pick = by_free[f].back(); by_free[f].pop_back(); }
        if (pick < 0) { bins.emplace_back(); pick = static_cast<int>(bins.size()) - 1; }
        bins[pick].rows += n[b];
        bins[pick].jets.push_back(b);
        by_free[128 - bins[pick].rows].push_back(pick);
    }
    p->meta.clear();
    p->row_slot.clear();
    for (const Bin& bin : bins) {
        TfTileMeta m{};
        std::vector<int> slots;
        for (int b : bin.jets) {
            const int beg = static_cast<int>(slots.size());
            for (int d = 0; d < D; ++d)
                if (mask[static_cast<size_t>(b) * D + d] != 0) slots.push_back(b * D + d);
            const int end = static_cast<int>(slots.size());
            for (int r = beg; r < end; ++r) {
                m.seg_beg[r] = static_cast<unsigned char>(beg);
                m.seg_end[r] = static_cast<unsigned char>(end);
                m.row_tb[r] = per_jet_time ? b : 0;
            }
        }
        m.nrows = static_cast<int>(slots.size());
        slots.resize(128, -1);
        p->meta.push_back(m);
        p->row_slot.insert(p->row_slot.end(), slots.begin(), slots.end());
    }
    if (!p->meta.empty()) {                              // pad with empty tiles to a whole number of clusters
        TfTileMeta empty{};
        while (p->meta.size() % cluster != 0) {
            p->meta.push_back(empty);
            p->row_slot.insert(p->row_slot.end(), 128, -1);
        }
    }
    p->n_plain = static_cast<int>(p->meta.size());
    // pair tiles: CTA 0 of the cluster takes the first ceil(n / 2) particles of the jet, CTA 1 the rest
    for (int b : big) {
        std::vector<int> s;
        for (int d = 0; d < D; ++d)
            if (mask[static_cast<size_t>(b) * D + d] != 0) s.push_back(b * D + d);
        const int h0 = (n[b] + 1) / 2, h1 = n[b] - h0;
        for (int half = 0; half < 2; ++half) {
            TfTileMeta m{};
            std::vector<int> slots(half == 0 ? s.begin() : s.begin() + h0, half == 0 ? s.begin() + h0 : s.end());
            m.nrows = half == 0 ? h0 : h1;
            m.pad[0] = half == 0 ? h1 : h0;
            for (int r = 0; r < m.nrows; ++r) {
                m.seg_beg[r] = 0;
                m.seg_end[r] = static_cast<unsigned char>(m.nrows);
                m.row_tb[r] = per_jet_time ? b : 0;
            }
            slots.resize(128, -1);
            p->meta.push_back(m);
            p->row_slot.insert(p->row_slot.end(), slots.begin(), slots.end());
        }
    }
    p->n_pair = 2 * static_cast<int>(big.size());
}

int ensure_ws(TfTileModel* m, int tiles, int tb) {
    if (tiles <= m->tile_cap && tb <= m->tb_cap) return 0;
    // grow with 25 % headroom: successive batches of a run differ by a few tiles, and every growth is a device-wide
    // synchronise + free + allocate of ~130 KB per tile (measured: up to 0.5 s on a 4096-jet batch)
    const int tc = tiles > m->tile_cap ? std::max(tiles + tiles / 4, 1) : m->tile_cap, bc = std::max(m->tb_cap, std::max(tb, 1));
    if (m->ws) { MMF_CUDA_OK(cudaDeviceSynchronize()); MMF_CUDA_OK(cudaFree(m->ws)); m->ws = nullptr; }
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    const size_t rows = static_cast<size_t>(tc) * 128;
    const size_t o_meta = take(static_cast<size_t>(tc) * sizeof(TfTileMeta)), o_xs = take(rows * 3 * 4), o_ks = take(rows * 4),
                 o_slot = take(rows * 4), o_skip = take(rows * 256 * 4), o_temb = take(static_cast<size_t>(bc) * 512 * 4),
                 o_th = take(static_cast<size_t>(bc) * 2 * 4);
    MMF_CUDA_OK(cudaMalloc(&m->ws, off));
    m->tile_cap = tc; m->tb_cap = bc;
    m->d_meta = reinterpret_cast<TfTileMeta*>(m->ws + o_meta);
    m->d_xs0 = reinterpret_cast<float*>(m->ws + o_xs);
    m->d_ks0 = reinterpret_cast<int*>(m->ws + o_ks);
    m->d_row_slot = reinterpret_cast<int*>(m->ws + o_slot);
    m->d_skip = reinterpret_cast<float*>(m->ws + o_skip);
    m->d_temb = reinterpret_cast<float*>(m->ws + o_temb);
    m->d_thermo = reinterpret_cast<float*>(m->ws + o_th);
    return 0;
}

}  // namespace

int tftile_prepare(TfTileModel* m, const TfRunArgs& r, std::vector<unsigned char>* handled, cudaStream_t s) {
    const MmfModelDesc& d = m->desc;
    const bool pf = d.arch == MMF_ARCH_PARTICLEFORMER;
    TilePlan plan;
    plan_tiles(r.mask_host, r.B, r.D, r.per_jet_time, m->cluster, &plan, handled);
    const int tiles = static_cast<int>(plan.meta.size());
    m->pending_tiles = plan.n_plain;
    m->pending_pair_tiles = plan.n_pair;
    if (tiles == 0) return 0;
    MMF_TRY_RC(ensure_ws(m, tiles, r.n_times));
    // time tables: sin/cos features (reference utils/models.py:62-75) and, for ParticleFormer, time_expand(temb)
    std::vector<float> temb(static_cast<size_t>(r.n_times) * 512, 0.f), thermo(static_cast<size_t>(r.n_times) * 2, 0.f);
    for (int i = 0; i < r.n_times; ++i) {
        float* row = &temb[static_cast<size_t>(i) * 512];
        if (pf) {
            sincos_row(r.times[i], 128, row);
            memcpy(row + 128, row, 128 * sizeof(float));
            for (int o = 0; o < 256; ++o) {
                float acc = 0.f;
                const float* wrow = &m->time_expand_w[static_cast<size_t>(o) * 128];
                for (int j = 0; j < 128; ++j) acc += wrow[j] * row[j];
                row[256 + o] = acc + m->time_expand_b[o];
            }
        } else {
            sincos_row(r.times[i], 256, row);
        }
        if (r.opts) det_thermostat(r.times[i], r.opts->beta, d.vocab_size, &thermo[i * 2], &thermo[i * 2 + 1]);
    }
    MMF_TRY_RC(m->stage.begin(temb.size() * 4 + thermo.size() * 4 + plan.meta.size() * sizeof(TfTileMeta) + plan.row_slot.size() * 4 + 64));
    MMF_TRY_RC(m->stage.push(m->d_temb, temb.data(), temb.size() * 4, s));
    MMF_TRY_RC(m->stage.push(m->d_thermo, thermo.data(), thermo.size() * 4, s));
    MMF_TRY_RC(m->stage.push(m->d_meta, plan.meta.data(), plan.meta.size() * sizeof(TfTileMeta), s));
    MMF_TRY_RC(m->stage.push(m->d_row_slot, plan.row_slot.data(), plan.row_slot.size() * 4, s));
    MMF_TRY_RC(m->stage.end(s));                         // no host synchronisation: the tables travel through pinned memory
    MMF_TRY_RC(launch_pack(r.x0, r.k0, m->d_row_slot, tiles * 128, d.vocab_size, m->d_xs0, m->d_ks0, r.err_flag, s));
    m->launches += 1;

    TfLaunch a{};
    a.arch = d.arch; a.n_stream = pf ? d.n_layer : 0; a.n_main = pf ? d.n_layer_fused : d.n_layer; a.vocab = d.vocab_size;
    a.optab = m->optab.get(); a.prodtab = m->prodtab.get(); a.n_ops = m->n_ops; a.n_prod = m->n_prod; a.n_blobs = m->n_blobs; a.wstream = m->d_wstream; a.params = m->d_params; a.meta = m->d_meta; a.tile0 = 0;
    a.xs0 = m->d_xs0; a.ks0 = m->d_ks0; a.row_slot = m->d_row_slot; a.skip = m->d_skip; a.temb = m->d_temb;
    a.per_jet_time = r.per_jet_time ? 1 : 0; a.nsteps = r.nsteps; a.softmax_nomax = m->softmax_nomax;
    if (r.opts) {
        a.st.sp = StepParams{r.opts->temperature, r.dt, r.opts->beta, r.opts->top_p, r.opts->top_k, d.vocab_size};
        a.st.seed = r.opts->seed;
        a.st.slot0 = r.opts->first_global_jet * static_cast<uint64_t>(r.D);
        a.st.argmax_last = r.opts->use_final_max_rates ? 1 : 0;
    }
    a.st.u = r.u; a.st.forced = r.forced; a.st.thermo = m->d_thermo; a.st.rates_out = r.rates_out; a.st.err_flag = r.err_flag;
    a.st.slots = static_cast<long long>(r.B) * r.D;
    a.x_out = r.x_out; a.k_out = r.k_out; a.vt_out = r.vt_out; a.logits_out = r.logits_out;
    m->pending = a;
    return 0;
}

int tftile_launch(TfTileModel* m, cudaStream_t s) {
    const int tiles = m->pending_tiles, pair_tiles = m->pending_pair_tiles;
    if (tiles + pair_tiles == 0) return 0;
    TfLaunch a = m->pending;
    const char* trace_path = getenv("MMF_TRACE");
    unsigned long long* d_trace = nullptr;
    if (trace_path) {
        MMF_CUDA_OK(cudaMalloc(&d_trace, 2048 * 8));
        MMF_CUDA_OK(cudaMemsetAsync(d_trace, 0, 2048 * 8, s));
        a.trace = d_trace;
    }
    // Plain tiles and pair tiles are two launches of the same program (different operand layout).  With both kinds in a
    // batch the pair tiles go to a side stream, so a few large jets fill SMs the plain tiles leave free instead of
    // waiting behind them (fork / join with events; `s` never runs ahead of either launch).
    const bool fork = tiles > 0 && pair_tiles > 0 && !d_trace && getenv("MMF_NO_OVERLAP") == nullptr;
    cudaStream_t sp = s;
    if (fork) {
        if (!m->side) {
            MMF_CUDA_OK(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
            MMF_CUDA_OK(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
            MMF_CUDA_OK(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
        }
        MMF_CUDA_OK(cudaEventRecord(m->ev_fork, s));
        MMF_CUDA_OK(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
        sp = m->side;
    }
    if (pair_tiles > 0) {
        TfLaunch ap = a;
        ap.optab = m->optab_pair.get(); ap.n_ops = m->n_ops_pair; ap.tile0 = tiles;
        MMF_TRY_RC(d_trace && tiles == 0 ? launch_tf_tiles_trace(ap, pair_tiles, 2, true, sp) : launch_tf_tiles(ap, pair_tiles, 2, true, sp));
        m->launches += 1;
        if (fork) MMF_CUDA_OK(cudaEventRecord(m->ev_join, sp));
    }
    if (tiles > 0) {
        MMF_TRY_RC(d_trace ? launch_tf_tiles_trace(a, tiles, m->cluster, false, s) : launch_tf_tiles(a, tiles, m->cluster, false, s));
        m->launches += 1;
    }
    if (fork) MMF_CUDA_OK(cudaStreamWaitEvent(s, m->ev_join, 0));
    if (d_trace) {
        std::vector<unsigned long long> hbuf(2048);
        if (cudaMemcpyAsync(hbuf.data(), d_trace, 2048 * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
            tf_tiles_dump_timeouts();
            MMF_REQUIRE(false, "tile kernel: traced launch failed (stderr lists the barrier waits that timed out)");
        }
        MMF_CUDA_OK(cudaStreamSynchronize(s));
        cudaFree(d_trace);
        if (FILE* f = fopen(trace_path, "w")) {
            const unsigned long long kClk = 0x00ffffffffffffffull;
            static const char* kTag[] = {"", " [before a wait]", " [done0 arrived]", " [done1 arrived]", " [done2 arrived]", " [done3 arrived]", " [kvfree arrived]"};
            for (int st = 0; st < 2; ++st)
                for (int i = 0; i < 512 && hbuf[st * 512 + i]; ++i) {
                    const unsigned long long c = hbuf[st * 512 + i] & kClk, c0 = hbuf[st * 512] & kClk, cp = i ? hbuf[st * 512 + i - 1] & kClk : c;
                    const unsigned tag = static_cast<unsigned>(hbuf[st * 512 + i] >> 56);
                    fprintf(f, "step %d mark %3d  +%llu cycles (total %llu)%s\n", st, i, c - cp, c - c0, tag < 7 ? kTag[tag] : "");
                }
            // when each of the first 128 MMA ops of timestep 1 was issued, relative to the step start
            for (int i = 0; i < 128; ++i)
                if (hbuf[1024 + i]) fprintf(f, "op %3d issued %lld\n", i, (long long)(hbuf[1024 + i] - (hbuf[512] & kClk)));
            fclose(f);
        }
    }
    return 0;
}

}  // namespace mmf
