// plan.h — host-side execution plan of the gate executor (no CUDA dependency).
//
// Pipeline:  qcb_op[] (reference vocabulary, qubit 0 = MSB)
//              -> lower()      Gate[]   (logical index-bit space, bit = n-1-qubit; quirks applied)
//              -> schedule()   Plan     (stages of fused shared-memory tile sweeps + qubit exchanges)
//              -> encode       flat 64-bit word stream = the "program" the tile kernel interprets
//
// The same word stream is (a) uploaded to the GPU and interpreted by k_tile_stage (kernels.cu),
// (b) returned by qcb_plan_serialize() so the CPU test-suite can check the scheduler and the encoding
// against the oracle with the host emulator in tests/emu/ (test infrastructure, not a product path).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/qcb200.h"

namespace qcb {

struct cplx { double re, im; };

// ---- lowered gates (logical bit space) ----
enum GateKind : int {
  G_MAT1 = 0,    // dense 2x2 on target bit t0, applied where (idx & cmask) == cmask
  G_MAT2 = 1,    // dense 4x4 on bits (t1 = more significant basis bit, t0 = less significant)
  G_SWAPP = 2,   // exchange amplitudes whose bits t0,t1 differ, multiplied by phase m[0]; control cmask
  G_DMASK = 3,   // diagonal: amp *= m[0] where (idx & dmask) == dval
  G_DTAB1 = 4,   // diagonal: where (idx & cmask) == cmask: amp *= (bit t0 ? m[1] : m[0])
  G_DPOP1 = 5,   // diagonal: amp *= m[0] where popcount(idx & dmask) == 1
  G_REFLECT = 6, // Grover diffusion 2|s><s| - I  (needs the global mean: reduction + affine pass)
  G_DENSE = 7,   // fused dense 2^r x 2^r block on ALL slot bits of its round (built by the round fuser)
};

struct Gate {
  int kind = G_MAT1;
  int t0 = -1, t1 = -1;        // target bits (non-diagonal action); for G_DTAB1 t0 is the table bit
  uint64_t cmask = 0;          // control bits (all must be 1)
  uint64_t dmask = 0, dval = 0;
  cplx m[16] = {};
  std::vector<cplx> dense;     // G_DENSE: row-major 2^r x 2^r in slot space
  int fused = 0;               // G_DENSE: number of gates folded in
  double frac = 1.0;           // fraction of the state a stand-alone application touches
  int src_op = -1;             // index of the originating qcb_op
  int uid = -1;                // scheduler: index of the gate in Plan::gates (which gates a stage really absorbed)

  uint64_t target_mask() const;   // bits acted on non-diagonally
  uint64_t diag_mask() const;     // bits only read (controls, diagonal operands)
};

// ---- device program encoding ----
constexpr int OP_WORDS = 16;              // one op slot = 16 x 64-bit words
constexpr int MAX_TILE_BITS = 13;
constexpr int MAX_SLOT_BITS = 3;          // register-resident bits per round
constexpr int MAX_RUNS = 16;

// D_MAT1: general complex 2x2; D_MAT1R: real 2x2 (H, RY, ...); D_MAT1RI: real diagonal + imaginary off-diagonal
// (RX, Y); D_PERMX: pair exchange (X, CNOT, Toffoli); D_DNEG: masked sign flip (Z, CZ, phase oracle)
enum DevOpKind : uint32_t { D_MAT1 = 0, D_MAT2 = 1, D_SWAPP = 2, D_DMASK = 3, D_DNEG = 4, D_DPOP1 = 5, D_AFFINE = 6,
                            D_MAT1R = 7, D_MAT1RI = 8, D_PERMX = 9, D_DENSE = 10 };

// Descriptor words (all uint64) — see encode_stage() for the authoritative writer.
//   StageDesc  (STAGE_WORDS words):
//     [0] n_local  [1] m  [2] L  [3] n_rounds  [4] n_runs  [5] ext_hi_base (rank << (n_local-m))
//     [6] skip_mask [7] skip_val  (tile skipped unless (ext_hi & skip_mask) == skip_val)
//     [8..8+MAX_TILE_BITS)  physical position of tile bit k (k < m), ascending; first L are 0..L-1
//     [24..24+MAX_RUNS)     run_start | run_len << 8 : contiguous runs of non-tile positions (ascending)
//     [40] total words of this stage (desc + rounds + ops)  [41] flags  [43] layout_c (tile_core.h: swz; > 0 = TMA mover)
//   Op slot (OP_WORDS words; MAT2 uses 3 slots):
//     [0] kind | j0<<8 | j1<<16 | n_slots<<24 | sel<<32   (sel: bitmap over slot patterns, see tile_core.h)
//     [1] loc_mask | loc_val<<32   (tile-local non-slot bits)      [2] hi_mask  [3] hi_val  (tile-id + rank bits)
//     [4..] payload: 2x2 / 4x4 matrix, phase, popcount mask, or device-value index
//   RoundDesc  (ROUND_WORDS words) x n_rounds, then op slots.
//     [0] r  [1] n_op_slots  [2] op word offset (from stage start)  [3] n_lane
//     [4..7)  slot_pos[3] (tile-local, ascending)   [7..10) lane_pos[3]
//     [10] n_ins  [11..17) ins_pos[6] ascending (slot ∪ lane positions)
//     [17] kind (0 = interpreter round, 1 = tensor-core round as a 16x16 real block, 2 = tensor-core round in the
//          three-product form, tile_core.h: K3Ctx, 3 = two three-product rounds on disjoint slot triples in one pass,
//          tile_core.h "paired rounds")   [18] n_grp_bits   [19..29) grp_pos[10]
//     [29] k (condition bits)  [30..34) cond_pos[4] (ext positions)  [34] j_load (kinds 2, 3: kmap)  [35] j_store (kinds 2, 3: mmap)
//     [36] kind 3: mmap2   [37] / [38] far-phase tables of the first / second block: word offset of the table applied AFTER the
//     block (row scaling) | offset of the one applied BEFORE it (column scaling) << 32, 0 = none
//     [39] entries: after-1 | after-2 << 8 | before-1 << 16 | before-2 << 24
//     tensor-core rounds: [2] = word offset of the A-fragment matrices (2^k * 256 / 192 / 384 doubles, after all descriptors)
//   StageDesc [42] = number of leading words (descriptors + interpreter op slots) that the kernel copies to smem
constexpr int STAGE_WORDS = 48;
constexpr int ROUND_WORDS = 40;
constexpr uint64_t FLAG_NEEDS_SUM = 1;     // epilogue: accumulate sum of amplitudes (for the next REFLECT)
constexpr uint64_t FLAG_DIRECT_STORE = 2;  // the last round writes its results to global memory itself (tile_core.h: T_FLAG_DIRECT_STORE)

// S_GROVER: one streaming pass a' = alpha * a + beta (the Grover diffusion, coefficients from the preceding sum) followed by
// the sign flips of the phase oracles that come next, and - when another diffusion follows - the sum of the result, so that
// a Grover iteration is ONE 32-byte-per-amplitude pass (kernels.cu: k_grover_step)
enum StageKind : int { S_TILE = 0, S_EXCHANGE = 1, S_SUM = 2, S_GROVER = 3 };
constexpr int MAX_GROVER_MARKED = 8;

constexpr int MAX_COND_BITS = 4;          // outside-condition bits of a tensor-core round (2^k matrix variants)

// roles the lane-assignment search of a tensor-core round chose (plan.cpp: build_k3_round / build_k3_pair_round); a plan trace
// keeps them so that a replay skips the search
struct RoundLayout { int kmap[3], mmap[3], mmap2[3], lanes[3]; };

struct Round {
  const RoundLayout* layout_hint = nullptr;   // replay: take the roles from here instead of searching
  std::vector<int> slot_pos;        // tile-local positions held in registers
  std::vector<Gate> gates;          // gates with bits already translated to ext space (see Stage)
  // tensor-core ("dmma") round: the whole round is one of 2^k dense 16x16 real matrices, selected per group
  // by the values of k outside-condition bits (controls / diagonal operands that are not slot bits)
  bool dmma = false;
  std::vector<int> cond_pos;        // ext positions of the condition bits (variant index bit j <-> cond_pos[j])
  std::vector<int> grp_pos;         // tile-local position of group-index bit i (first 3 = lane bits)
  int j_load = 0, j_store = 0;      // slot index paired with k bit 1 (loads) / m bit 1 (stores): bank-conflict control
  bool k3 = false;                  // round kind 2 (tile_core.h: K3Ctx): three 8x8 real matrices, six m8n8k4 steps per batch
  int kmap[3] = {0, 1, 2}, mmap[3] = {0, 1, 2};   // k3: slot index carried by bit b of the k-index (loads) / m-index (stores)
  std::vector<double> frag;         // kind 1: 2^k * 256 doubles in mma.m16n8k16 A-fragment order [variant][reg][lane];
                                    // kind 2: 2^k * 192 doubles [variant][P0 P1 N0 N1 R0 R1][lane]
  std::vector<int> uids;            // Gate::uid of the gates the scheduler put into this round, in order (plan traces)
  uint64_t slot_mask = 0;           // the slot bits the scheduler chose (before a tensor-core round pads them to 3)
  // paired round (round kind 3, tile_core.h "paired rounds"): a second dense block, on the disjoint slot triple
  // grp_pos[0..2], applied to the first block's results straight from registers - one pass over the tile for two rounds
  bool pair = false;
  std::vector<Gate> gates2;         // gates of the second block (ext space), applied after `gates`
  std::vector<int> uids2;
  uint64_t slot_mask2 = 0;          // slot bits the scheduler chose for the second block (tile-local, before padding)
  int mmap2[3] = {0, 1, 2};         // index into grp_pos[0..2] of the bit carried by bit b of the second block's m-index
  int dense_rounds() const { return pair ? 2 : 1; }
  // far phases (tile_core.h "far phases"): diagonal two-bit gates with one operand on a slot and the other OUTSIDE the tile are
  // not condition bits - the product of their phases is a constant of the tile, applied by every lane to its A fragments.
  // Per entry five words: ext position of the far bit minus m, then gamma, phi_0, phi_1, phi_2 (doubles) - see build_far_table.
  std::vector<uint64_t> far, far2;        // after the first / second block: row scaling
  std::vector<uint64_t> farpre, farpre2;  // before the first / second block (no earlier non-diagonal gate on the slot): column scaling
};

struct Stage {
  int kind = S_TILE;
  // S_TILE
  int m = 0, L = 0;
  int layout_c = 0;                 // layout parameter of the shared tile (tile_core.h: swz): 0 = full XOR fold (LSU mover);
                                    // TMA mover: number of contiguous low tile bits = run of one tensor copy (3..11)
  std::vector<int> tile_pos;        // physical positions of the tile bits, ascending
  uint64_t skip_mask = 0, skip_val = 0;
  uint64_t flags = 0;
  std::vector<Round> rounds;
  std::vector<int> src_gates;       // indices into the lowered gate list (for tests / stats)
  std::vector<int> absorbed;        // scheduler scratch: Gate::uid of every gate the formed rounds hold
  double sweep_fraction = 1.0;      // fraction of tiles actually visited
  // S_EXCHANGE: swap global physical bit gbit with local physical bit lbit
  int gbit = -1, lbit = -1;
  // S_GROVER: physical LOCAL indices of the marked states that live on this rank; needs_sum = a diffusion follows
  std::vector<uint64_t> marked;
  bool needs_sum = false;
};

struct Config {
  int n_total = 0, n_local = 0;
  int rank = 0, world = 1;
  int tile_bits = 12, low_bits = 4;
  int fusion = 1, strict = 1;
  int max_stage_cost = 0;
  int max_stage_rounds = 0;
  int dense_mma = 1;           // rounds as dense 8x8 complex blocks on the fp64 tensor cores
  int direct_store = 0;        // last tensor-core round of a sweep stores straight to HBM (no STS + mover read-back); measured
                               // neutral to slightly slower (profiles/r2c_ab.log: 5423 vs 5521 gates/s), hence off (QCB_DIRECT_STORE=1)
  int mma_form = 0;            // 0 = three-product form (six m8n8k4 steps per batch, 16-byte shared accesses);
                               // 1 = 16x16 real block (m16n8k16 = eight steps, 8-byte accesses), the round-1 kernel
  int round_yield_pct = 50;    // end a stage early when the next round would absorb less than this % of the stage's average round
  int window_search = 1;       // stage builder also tries contiguous tile windows and keeps the best yield
  int pair_rounds = 1;         // consecutive tensor-core rounds on disjoint slot triples share one pass over the tile (round kind 3)
  int pair_search = 1;         // the partner search is run for this many of the best candidate rounds
  int pair_eff_pct = 170;      // cost of a paired pass in % of a single round (measured on B200: 3.39 ms vs 2.0 ms at 30 qubits):
                               // a pair is formed when its gates per unit of cost beat the best single round's
  int pair_cost_q = 7;         // cost of a paired pass in quarter rounds (a single round = 4) against the stage's round budget
  int far_phase = 1;           // diagonal two-bit gates with a far operand (outside the tile) ride as tile-constant phases
  int plan_portfolio = 1;      // large circuits: schedule under a handful of budget settings, keep the plan the cost model prefers
  int thin_defer = 12;         // multi-GPU: a stage with fewer gates than this is not run while gates wait for an exchange
                               // (its gates ride along in the fuller sweeps after the exchange)
  int tma = 0;                 // tiles move by TMA tensor copies (layout follows the hardware 128-byte swizzle)
  int threads = 256;
};

struct Plan {
  Config cfg;
  std::vector<Gate> gates;                 // lowered
  std::vector<Stage> stages;
  std::vector<int> perm_out;               // logical bit -> physical bit after the plan ran
  std::vector<uint64_t> words;             // encoded program: [hdr][stage...]
  std::vector<uint64_t> stage_offsets;     // word offset of each stage in `words`
  uint64_t n_rounds = 0, n_exchanges = 0;
  double algorithmic_bytes = 0, unfused_bytes = 0;
  std::string error;
  // Support of the state: LOGICAL bits that can be 1 in the index of a non-zero amplitude (~0 = nothing known).  A state
  // straight after qcb_set_zero has support 0; every non-diagonal target of an executed gate joins it.  A tile sweep over a state
  // with known support visits only the tiles whose id bits outside the support are 0 (Stage::skip_mask).  EXPERIMENTAL: the
  // handle passes anything but ~0 only with QCB_ZERO_SKIP=1 (sim.cu); verified on the host emulator, not yet on hardware.
  uint64_t support_in = ~0ULL, support_out = ~0ULL;
};

// ---- plan traces: the scheduler's decisions for one circuit STRUCTURE (which gates share a sweep / a round, tile and slot
// bits, exchanges), recorded once and replayed for every later circuit with the same structure and different angles
// (variational loops: VQE / QAOA evaluate the same ansatz thousands of times).  Replay skips the tile-window and
// slot-triple searches; the matrices are rebuilt from the new gates.  `key` holds everything a decision depends on.
struct StageTrace {
  int kind = S_TILE;
  bool lead = false;                         // S_TILE opened by the affine pass of a Grover diffusion
  int lead_uid = -1;                         // S_SUM / lead: the G_REFLECT gate
  uint64_t tile_bits = 0;                    // tile choice handed to the stage builder
  std::vector<int> taken;                    // gates handed to the stage builder (uids, in order)
  std::vector<std::vector<int>> round_uids;  // per formed round
  std::vector<uint64_t> round_slots;
  std::vector<uint8_t> round_pair;           // per formed round: 1 = this round and the next one share a pass (round kind 3)
  std::vector<RoundLayout> layouts;          // per PASS of three-product rounds: the lane roles (empty entries for other kinds)
  int gbit = -1, lbit = -1;                  // S_EXCHANGE
  bool needs_sum = false;                    // S_GROVER (taken = the diffusion and the oracles it absorbs)
};
struct PlanTrace {
  size_t words_hint = 0;                     // size of the encoded program (reserve on replay)
  std::vector<uint64_t> key;
  std::vector<StageTrace> stages;
};
// Everything the scheduler's decisions depend on: configuration, incoming bit permutation, and per gate its kind, bits,
// cost class and sign-flip flag (never the angles themselves).
void plan_structure_key(const Config& cfg, const std::vector<Gate>& gates, const std::vector<int>& perm_in, std::vector<uint64_t>& key,
                        uint64_t support_in = ~0ULL);

// qcb_op[] -> Gate[] ; returns QCB_OK or an error code with `err` set.
int lower_ops(const Config& cfg, const qcb_op* ops, uint64_t n_ops, std::vector<Gate>& out, std::string& err);

// Receives every stage as soon as the scheduler has finished and encoded it (Plan::words / stage_offsets are valid up
// to and including that stage), so that execution can start while later stages are still being planned.
struct StageSink {
  virtual ~StageSink() {}
  virtual int on_stage(Plan& plan, size_t stage_index) = 0;     // QCB_OK or an error code (aborts scheduling)
};

// Gate[] -> stages (+ encoded program).  perm_in: logical->physical bit map (identity when empty).
// record != nullptr: the decisions are written to *record (its key is filled in).  replay != nullptr: the decisions are
// taken from *replay instead of searched (the caller has checked that the keys match).
int schedule(Plan& plan, const std::vector<int>& perm_in, StageSink* sink = nullptr, PlanTrace* record = nullptr,
             const PlanTrace* replay = nullptr);

// Standard 2x2 matrices of the reference (domain/gate.clj:38-283)
void gate_matrix(int kind, double angle, cplx out[4]);

Config config_from(const qcb_config& c);

}  // namespace qcb
