// ParticleFormer / FusedParticleFormer sampler as ONE persistent kernel: a CTA owns a 128-row tile of whole
// jets (each <= 128 particles) for all N timesteps; activations never leave the SM.
//   residual stream [128 x 256] fp32           TMEM columns [0,256)
//   scratch accumulators                        TMEM columns [256,512)
//   GEMM operands (bf16, SWIZZLE_128B)          shared memory arena, rewritten by the epilogue warps
//   weights                                     bf16 tiles streamed from L2 in consumption order (1-D bulk copies)
//   per-stage fp32 vectors (biases, LN affine)  "parameter blobs" streamed into a double buffer
// The tensor-core work of one timestep is a flat table of k-tile operations (TfOp) built on the host; the MMA
// issuer and the weight producer interpret it, the epilogue warps run the matching hand-written program.
// reference: networks/ParticleTransformers.py:62-122, :177-210; networks/attention.py:23-26, 53-74;
//            model/solvers.py:22-60 (the step fused after the heads).
#pragma once
#include "mmf_host.h"
#include "mmf_internal.h"

namespace mmf {

constexpr int kTfParamFloats = 2304;            // floats per parameter blob slot

// float offsets inside the parameter blobs (one blob per stage of the per-timestep program, <= kTfParamFloats each)
namespace tfp {
// embedding stage A: continuous branch
constexpr int EA_W0 = 0;                         // [256][4] = wxe.0 weight (3) and bias
constexpr int EA_BXE2 = 1024;                    // [128] wxe.2 bias
constexpr int EA_LN1X_G = 1152, EA_LN1X_B = 1280;  // [128]
// embedding stage B: discrete branch + first LayerNorm
constexpr int EB_YTAB = 0;                       // [V][128], V <= 12: LN_ln1y(wye.2(GELU(wye.0[k])))
constexpr int EB_LNN_G = 1536, EB_LNN_B = 1792;  // [256] LayerNorm ln1 of the first block
// main block (C = 256), attention stage
constexpr int BA_BQKV = 0;                       // [768] q | k | v
constexpr int BA_QG = 768, BA_QB = 832, BA_KG = 896, BA_KB = 960;    // [64]
constexpr int BA_BPROJ = 1024;                   // [256]
constexpr int BA_LN2G = 1280, BA_LN2B = 1536;    // [256]
// main block, MLP stage
constexpr int BM_BFC = 0;                        // [512]
constexpr int BM_BP2 = 512;                      // [256]
constexpr int BM_LNN_G = 768, BM_LNN_B = 1024;   // [256] next LayerNorm (ln1 of the next block, or the final one)
// stream block (two groups of C = 128), attention stage; group g starts at g * SA_GROUP
constexpr int SA_GROUP = 896;
constexpr int SA_BQKV = 0;                       // [384] q | k | v
constexpr int SA_QG = 384, SA_QB = 416, SA_KG = 448, SA_KB = 480;    // [32]
constexpr int SA_BPROJ = 512;                    // [128]
constexpr int SA_LN2G = 640, SA_LN2B = 768;      // [128]
// stream block, MLP stage; group g starts at g * SM_GROUP
constexpr int SM_GROUP = 640;
constexpr int SM_BFC = 0;                        // [512]
constexpr int SM_BP2 = 512;                      // [128]
constexpr int SM_LNN_G = 1280, SM_LNN_B = 1536;  // [256] next LayerNorm over both groups (ln1 of next block; last: ln2_x|ln2_y)
constexpr int SM_LN2ND_G = 1792, SM_LN2ND_B = 2048;  // [256] last stream block only: ln1 of the first main block
// head stages
constexpr int HX_BIAS = 0;                       // [512] head_x.0 bias
constexpr int HX_W2 = 512;                       // [3][512] head_x.2 weight
constexpr int HX_B2 = 2048;                      // [3]
constexpr int HY_BIAS = 0;                       // [128] head_y.0 bias, this quarter of the hidden units
constexpr int HY_W2 = 128;                       // [V][128] head_y.2 weight, this quarter
constexpr int HY_B2 = 128 + 12 * 128;            // [V] (first quarter only), V <= 12
}  // namespace tfp

// One tensor-core operation = `nkt` k-tiles (K = 64 each, or one K = 32 k-tile) accumulated into the same D columns.
// Everything the issuing warp needs is precomputed on the host: the low words of the shared-memory matrix
// descriptors (relative to the operand arena) and the instruction descriptor.
//   A advances by one 16 KB chunk per k-tile; B is the next `nkt` tiles of the weight ring, or an arena slice
//   advancing by 8 KB per k-tile (V^T halves of the P V product).
constexpr uint32_t kTfOpAcc = 1u, kTfOpWait = 2u, kTfOpRing = 4u, kTfOpHalfK = 8u, kTfOpBMn = 64u, kTfOpAttn = 128u;   // TfOp.flags
struct TfOp {
    uint32_t a_lo;          // (arena offset of A >> 4) | LBO field of the descriptor low word
    uint32_t b_lo;          // same for B when it lives in the arena (ignored for ring operands)
    uint32_t idesc;         // kind::f16 instruction descriptor (M = 128, N)
    uint16_t dcol;          // bits 0-9: TMEM column of D (relative to the allocation base); bits 10-15: kTfPair* hand-offs
    uint8_t nkt;            // bits 0-2: k-tiles (0: a pure hand-off op of a pair tile); bit 7: third bit of the completion signal (signal 4 = commit -> done[3])
    uint8_t flags;          // kTfOpAcc: the first MMA accumulates onto D; kTfOpWait: wait for the next "go" of the epilogue
                            // warps first; kTfOpRing: B from the weight ring; kTfOpHalfK: K = 32 (2 MMAs) instead of 64 (4);
                            // bits 4-5 (+ nkt bit 7): after the last k-tile 0 nothing, s = 1..4 commit -> done[s - 1];
                            // kTfOpAttn: issued by the attention issuer warp (S and P V: both operands in the arena) and handed
                            // off on its own barrier; every other op belongs to the weight-GEMM issuer, which alone reads the ring;
                            // kTfOpBMn: B is MN-major ([K rows][N] with N contiguous: the V operand of P V as the epilogue
                            // writes it, no transpose) - each K = 16 step advances B by 16 rows of 128 bytes
};
static_assert(sizeof(TfOp) == 16, "TfOp is read with one 128-bit load");

// Pair tiles: hand-offs of the K / V exchange between the two CTAs of a cluster, carried in the high bits of TfOp.dcol
// (the TMEM column needs 9 bits).  All of them are executed by the attention-issuer warp around the op they sit on.
constexpr uint32_t kTfDcolMask = 0x3ffu;
constexpr uint32_t kTfPairShift = 10;
constexpr uint32_t kTfPairSendK = 1u;        // before the op: (partner done with its K rows of the previous exchange) copy my K rows to it
constexpr uint32_t kTfPairSendVWaitK = 2u;   // before the op: (partner done with its V rows) copy my V rows to it; wait for its K rows
constexpr uint32_t kTfPairFreeK = 4u;        // after the op: commit -> partner's kfree (my MMAs have read K for the last time)
constexpr uint32_t kTfPairWaitV = 8u;        // before the op: wait for the partner's V rows
constexpr uint32_t kTfPairFreeV = 16u;       // after the op: commit -> partner's vfree

constexpr int kTfMaxOps = 1024;
struct TfOpTable {          // MMA ops of one timestep; passed to the kernel BY VALUE (constant bank -> uniform registers)
    TfOp ops[kTfMaxOps];
};
// Weight tiles of one timestep in consumption order.  The ring is a 64 KB byte range; a tile of `kb` KB ([n rows][64] bf16,
// n = 8 kb) is copied to offset `dst` KB.  Placement is planned on the host (tiles never wrap; the plan is the same every
// timestep): before tile g (global index over the launch) is written, all tiles up to g - dep must have been consumed.
// Tile g uses the barrier pair g % kTfRingBars.
//   x = offset in the weight stream / 128 | kb << 24        y = dst | dep << 8
constexpr int kTfRingBytes = 65536;
constexpr int kTfRingBars = 16;
struct TfProdTable {
    uint2 e[kTfMaxOps];
};

struct TfTileMeta {
    int nrows;              // real rows (<= 128; <= 80 in a pair tile)
    int pad[3];             // pad[0]: pair tiles - real rows of the partner CTA (the other half of the jet)
    unsigned char seg_beg[128];   // per row: first row of its jet
    unsigned char seg_end[128];   // per row: one past the last row of its jet (0,0 for padding rows)
    int row_tb[128];        // per row: row of the time table when time is per jet (forward API)
};

struct TfStepCfg {          // the hybrid step (sampler mode)
    StepParams sp;
    const float* u;         // (N, B*D, V) supplied uniforms or null -> Philox
    uint64_t seed, slot0;
    const unsigned char* forced;   // (N, B*D) teacher-forced tokens or null
    const float* thermo;    // [N][2] (w, coef) per timestep
    float* rates_out;       // padded (B*D, V) rates of the last step or null
    int argmax_last;        // use_final_max_rates
    int* err_flag;
    long long slots;        // B*D
};

struct TfLaunch {
    int arch;               // MMF_ARCH_PARTICLEFORMER / MMF_ARCH_FUSED_PARTICLEFORMER
    int n_stream, n_main;   // stream (2 x 128-wide) blocks and main (256-wide) blocks
    int vocab;
    const TfOpTable* optab; // HOST pointers: the tables of one timestep, copied into the kernel parameters at launch
    const TfProdTable* prodtab;
    int n_ops, n_prod;      // ops of the MMA issuer; weight tiles of the producers
    int n_blobs;            // parameter blobs per timestep (consumed in index order)
    const uint8_t* wstream; // weight tiles in op order
    const float* params;    // parameter blobs, kTfParamFloats apart
    const TfTileMeta* meta;
    int tile0;
    const float* xs0;       // [tiles*128][3] packed state
    const int* ks0;         // [tiles*128] packed tokens
    const int* row_slot;    // [tiles*128] packed row -> b*D + d, -1 padding
    float* skip;            // [tiles][256][128] fp32 skip stream, column-major per tile
    const float* temb;      // [*][512]: time embedding (256) | time_expand(temb) (256, ParticleFormer)
    int per_jet_time;
    int nsteps;
    int softmax_nomax;      // 1: the checkpoint's q / k LayerNorm parameters bound every score, the softmax skips the row maximum
    TfStepCfg st;
    float* x_out;           // padded (B,D,3) final state     (sampler)
    long long* k_out;       // padded (B,D) final tokens      (sampler)
    float* vt_out;          // padded (B,D,3) velocity        (forward API) or null
    float* logits_out;      // padded (B,D,V) logits          (forward API)
    unsigned long long* trace;
};

// Operand arena of a CTA (byte offsets; every region is 1024-byte aligned).  Two layouts:
//   plain tile  128 rows of whole jets, each <= 128 particles; keys = the tile's own 128 rows
//   PAIR tile   one jet of 129...160 particles split over the two CTAs of a cluster, <= 80 rows each.  The 160 keys of a CTA
//               are its own 80 rows followed by the partner's 80, whose K / V rows the partner's epilogue writes straight
//               into this CTA's shared memory (DSMEM).  Rows 80...127 of every A operand are don't-care, so A chunks
//               ([rows][64] bf16, SWIZZLE_128B) are laid 10 KB apart instead of 16 KB: the MMA reads M = 128 rows and
//               what it finds behind row 79 only reaches accumulator rows nobody looks at.  That makes room for K and V
//               of 160 keys and for a probability operand P that does not alias Q | K (the two heads of a 32-wide unit
//               take turns on the score columns, so head 1's Q and K must survive head 0's softmax).
template <bool PAIR>
struct TfLay {
    static constexpr uint32_t kChunk = PAIR ? 10240u : 16384u;      // stride of an A-operand chunk
    static constexpr uint32_t kRows = PAIR ? 80u : 128u;            // rows (and own keys) that exist in shared memory
    static constexpr uint32_t kKeys = PAIR ? 160u : 128u;           // keys a query row sees
    static constexpr uint32_t kKV = kKeys * 128u;                   // K / V: [keys][64] bf16
    static constexpr uint32_t oA = 0;                               // 4 chunks: LayerNorm output / head input
    static constexpr uint32_t oQ = 4 * kChunk;                      // Q of the current unit
    static constexpr uint32_t oK = oQ + kChunk;                     // K [keys][64]
    static constexpr uint32_t oVT = oK + kKV;                       // V [keys][64 d] (MN-major B operand of P V)
    static constexpr uint32_t oP = PAIR ? oVT + kKV : oQ;           // probabilities [rows][keys]; plain: aliases Q | K
    static constexpr uint32_t oO = PAIR ? oP + 3 * kChunk : oVT + kKV;   // attention output of the current unit
    static constexpr uint32_t oH0 = oQ, oH1 = oQ + 2 * kChunk;      // MLP hidden quarters [rows][128] (two chunks each)
    static constexpr uint32_t oRing = oO + kChunk;                  // weight ring
    static constexpr uint32_t kArena = oRing + 65536u;
    // scratch columns of TMEM (relative to the allocation base; the residual stream owns [0,256))
    static constexpr uint32_t cQkv64 = PAIR ? 320u : 256u;          // 64-wide units: q | k | v
    static constexpr uint32_t cS = PAIR ? 320u : 256u;              // scores (pair: one head at a time, 160 columns)
    static constexpr uint32_t cO64 = PAIR ? 256u : 448u;            // O of a 64-wide head
};
static_assert(TfLay<false>::oO == 114688u && TfLay<false>::kArena == 196608u, "plain arena layout");
static_assert(TfLay<true>::kArena == 198656u, "pair arena layout");

// host-side placement of the weight tiles in the 64 KB ring (tftile_model.cu); exported for tests as mmf_dbg_ring_plan
bool plan_weight_ring(const std::vector<int>& kb, std::vector<int>* dst_kb, std::vector<int>* dep);

int tf_tile_smem_bytes();
// n_tiles must be a multiple of `cluster` (1, 2 or 4): the CTAs of a cluster share each weight tile through TMA multicast
// pair = true: pair tiles (cluster must be 2; a.optab is the pair op table)
int launch_tf_tiles(const TfLaunch& a, int n_tiles, int cluster, bool pair, cudaStream_t stream);
int launch_tf_tiles_trace(const TfLaunch& a, int n_tiles, int cluster, bool pair, cudaStream_t stream);   // the same kernel with clock stamps
void tf_tiles_dump_timeouts();   // trace build: prints the barrier waits that timed out in a failed launch

// ---- host object (tftile_model.cu)
struct TfTileModel;
int tftile_create(const MmfModelDesc& d, WeightMap& wm, TfTileModel** out);
void tftile_destroy(TfTileModel* m);
int64_t tftile_launches(const TfTileModel* m);
// Jets of the batch that fit a tile (1 <= n <= 128) are handled here; `handled[b]` is set to 1 for them.
// x0/k0 and the outputs are device pointers in the padded layout; mask_host, times are host arrays.
struct TfRunArgs {
    const float* x0; const long long* k0; const int64_t* mask_host; int B, D;
    const float* times; int n_times; bool per_jet_time; int nsteps; float dt;
    const MmfStepOptions* opts; const float* u; const unsigned char* forced;
    float* x_out; long long* k_out; float* rates_out; float* vt_out; float* logits_out;
    int* err_flag;
};
// prepare: plan tiles, upload tables, pack the source state (reads x0/k0).  launch: run the kernel (writes the outputs of
// the handled jets).  The caller may zero the output buffers in between (they may alias the inputs).
int tftile_prepare(TfTileModel* m, const TfRunArgs& r, std::vector<unsigned char>* handled, cudaStream_t s);
int tftile_launch(TfTileModel* m, cudaStream_t s);

}  // namespace mmf
