// EPiC deep-set encoder on the B200: one persistent CTA per 128-row tile of WHOLE jets runs the entire
// N-step Euler sampler (reference networks/EPiC.py:38-62, model/CFM.py:133-154, model/solvers.py:139-143).
#pragma once
#include "mmf_host.h"
#include "mmf_internal.h"

namespace mmf {

constexpr int kEpicMaxJets = 8;         // jets per tile (pooled vectors live in shared memory)
constexpr int kEpicLayers = 5;          // the weight stream below is laid out for config.n_layer == 5
constexpr int kEpicTilesPerStep = 8 * (1 + 2 * kEpicLayers);   // 16 KB weight tiles consumed per timestep
constexpr int kEpicTbLd = 1800;         // floats per row of the time-bias table
// time-bias table row: [0,256) proj.mlp_local.0 | [256,512) proj.mlp_global.0 | [512+256 l, ..) fc_loc1 of layer l
//                      | [1792,1795) head

struct EpicTileMeta {
    int nrows;                          // real rows of this tile (<= 128)
    int njets;                          // jets of this tile (<= kEpicMaxJets; 1 for a pair tile)
    int pair;                           // 1: this CTA holds one half of a jet split over a 2-CTA cluster
    int pad;
    int jet_begin[kEpicMaxJets + 1];    // local row range of jet j
    int jet_ntot[kEpicMaxJets];         // particles of jet j (both halves for a pair tile): divisor of the mean
    int jet_tb[kEpicMaxJets];           // row of the time-bias table when time is per jet (forward API)
};

struct EpicParams {                     // device pointers into the packed checkpoint
    const uint8_t* wstream;             // kEpicTilesPerStep x 16 KB bf16 tiles [128 out][64 in], SWIZZLE_128B, in consumption order
    const float* a3;                    // [256][3]  proj.mlp_local.0[:, 256:] . wxe          (K = 3 fold)
    const float* b_loc2p;               // [256]     proj.mlp_local.2 bias
    const bf16* wg0t;                   // [512][256] proj.mlp_global.0[:, :512]^T
    const float *wg2p, *bg2p;           // [16][256], [16] proj.mlp_global.2
    const bf16* wg1t[kEpicLayers];      // [528][256] fc_glob1^T
    const float* bg1[kEpicLayers];      // [256]
    const float *wg2[kEpicLayers], *bg2[kEpicLayers];   // [16][256], [16] fc_glob2
    const float* wl1g[kEpicLayers];     // [256][16] fc_loc1[:, 512:528]
    const float* bl2[kEpicLayers];      // [256] fc_loc2 bias
    const float* wh_loc;                // [3][256] head[:, 256:512]
    const float* wh_glob;               // [3][16]  head[:, 512:528]
};

struct EpicLaunch {
    EpicParams p;
    const EpicTileMeta* meta;
    int tile0;                          // first tile of this launch
    const float* xs0;                   // [tiles*128][3] packed source state
    const int* row_slot;                // [tiles*128] packed row -> b*D + d (padding rows: -1)
    float* loc_skip;                    // [tiles*128][256] fp32 scratch (skip stream of the local features)
    const float* tbias;                 // [*][kEpicTbLd]
    int per_jet_time;                   // 1: table row = meta.jet_tb[j]; 0: table row = timestep
    int nsteps;
    float dt;
    float* x_out;                       // padded (B,D,3): final state (sampler)
    float* vt_out;                      // padded (B,D,3): velocity of one forward (forward API) or null
    unsigned long long* trace;          // optional [2 steps][64 marks] clock64 stamps of CTA 0 (MMF_TRACE=file), or null
};

struct EpicTimeFold {                   // inputs of the time-bias kernel
    const float* wt;                    // [7][256 k][256 o] transposed time columns of the 7 folded linears
    const float* cst;                   // [7][256] constants (biases, wxe bias fold)
    const float* wht;                   // [256 k][3] head[:, :256]^T
    const float* bh;                    // [3]
};

struct EpicModel;                       // host object: packed checkpoint + workspace (epic_model.cu)
int epic_create(const MmfModelDesc& d, WeightMap& wm, EpicModel** out);
void epic_destroy(EpicModel* m);
int64_t epic_launches(const EpicModel* m);
// mask_host (B,D) and the times are HOST arrays; x / outputs are device pointers in the padded (B,D,3) layout
int epic_forward(EpicModel* m, const float* x, const int64_t* mask_host, const float* t_host, int B, int D, float* vt_out,
                 cudaStream_t s);
int epic_generate(EpicModel* m, const float* x0, const int64_t* mask_host, int B, int D, const float* t_grid, int N, float dt,
                  float* x_out, cudaStream_t s);

int epic_smem_bytes();
// cluster == 2 launches pairs of CTAs (tiles 2i, 2i+1 hold the two halves of one jet)
int launch_epic_tiles(const EpicLaunch& a, int n_tiles, int cluster, cudaStream_t stream);
// tbias[tb] from temb[tb][256] for tb < n
int launch_epic_time_bias(const EpicTimeFold& f, const float* temb, int n, float* tbias, cudaStream_t stream);

}  // namespace mmf
