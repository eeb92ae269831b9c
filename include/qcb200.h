/*
 * qcb200.h — C ABI of libqcb200.so, the B200-native state-vector backend for QClojure.
 *
 * This is the drop-in boundary: the entry points are what a QClojure backend
 * (Clojure defrecord implementing `QuantumBackend`, application/backend.clj:72-112, reaching native
 * code through Java FFM on JDK 22+ or JNA) binds for the simulation hot path, and what the Python
 * test/bench harness binds through ctypes.  Each group below cites the reference interface it
 * replaces; paths are relative to /root/reference/src/org/soulspace/qclojure/.
 *
 * Conventions
 *   - every function returns int32 status: 0 = QCB_OK, < 0 = error class; the message is fetched
 *     with qcb_last_error().  No exceptions or longjmp cross this boundary.
 *   - all buffers are caller-allocated HOST memory unless the name says `_dev`; the library never
 *     keeps a host pointer after the call returns.
 *   - complex numbers are interleaved double[2] = (re, im); a state of n qubits is 2^n of them.
 *   - qubit indices use the REFERENCE's numbering: qubit 0 is the MOST significant bit of the
 *     amplitude index (domain/state.clj:114-162); the library translates to bit positions inside.
 *   - a handle is used by one thread at a time; different handles are independent (own stream).
 *   - multi-GPU: either one handle per process and rank (qcb_config.rank / world_size / nccl_unique_id, SPMD under
 *     torchrun; buffers then address the rank's slice), or ONE handle for all devices of the process (qcb_config.n_gpus;
 *     buffers address the whole state).
 *   - there is NO CPU fallback: qcb_create fails with QCB_ERR_CUDA when no CUDA device is usable.
 */
#ifndef QCB200_H
#define QCB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCB_ABI_VERSION 2

/* ---- status codes ---- */
#define QCB_OK              0
#define QCB_ERR_INVALID    -1   /* bad argument / malformed op (reference: ex-info / AssertionError) */
#define QCB_ERR_UNSUPPORTED -2  /* "Unknown gate type" (domain/circuit.clj:1072) and friends        */
#define QCB_ERR_CUDA       -3   /* CUDA runtime / no device                                         */
#define QCB_ERR_NOMEM      -4
#define QCB_ERR_NCCL       -5
#define QCB_ERR_STATE      -6   /* e.g. "State is not properly normalized" (domain/state.clj:900)   */
#define QCB_ERR_NOTFOUND   -7   /* unknown job id (job-status :not-found)                           */

typedef struct qcb_sim* qcb_handle;

/* ---- configuration (replaces the option maps of create-simulator, adapter/backend/ideal_simulator.clj:181-195) ---- */
typedef struct qcb_config {
  int32_t n_qubits;        /* total qubits of the state (all ranks together)                          */
  int32_t device;          /* CUDA device ordinal; -1 = current device                                */
  int32_t fusion;          /* 1 (default) = tile-fused execution; 0 = one sweep per gate (unfused)     */
  int32_t strict_parity;   /* 1 (default) = reproduce the reference literally: controlled gates apply
                              the transposed 2x2 (domain/gate.clj:473-483), SWAP/iSWAP qubit numbers
                              count from the LSB (gate.clj:768-778), :i and :cy are rejected like
                              circuit.clj:1072.  0 = textbook semantics + :i/:cy accepted.            */
  int32_t tile_bits;       /* 0 = default; log2 amplitudes of one shared-memory tile                  */
  int32_t low_bits;        /* 0 = default; number of always-resident low index bits (coalescing run)  */
  int32_t rank;            /* multi-GPU: this handle's rank in [0, world_size)                        */
  int32_t world_size;      /* power of two; 1 = single GPU                                            */
  const void* nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks (qcb_nccl_unique_id), or NULL */
  int32_t max_stage_cost;  /* 0 = default; scheduler knob: cost units one fused sweep may absorb       */
  int32_t max_stage_rounds;/* 0 = default; scheduler knob: shared-memory rounds one fused sweep may hold */
  int32_t dense_mma;       /* 0 = default (on); 1 = on: rounds run as dense 8x8 complex blocks on the fp64 tensor
                              cores (DMMA) in the three-product form (six DMMA.8x8x4 per 8 groups); 2 = off: register-
                              resident op interpreter only; 3 = on, 16x16 real form (eight DMMA.8x8x4 per 8 groups)     */
  int32_t tile_mover;      /* 0 = default; 1 = tiles move between HBM and shared memory with cp.async / st.global
                              (16 bytes per thread, arbitrary swizzle: conflict-free rounds); 2 = TMA tensor copies,
                              one per contiguous run, with the hardware 128-byte swizzle                            */
  /* ---- single-process multi-GPU (ABI 2): ONE handle owns the whole sharded state.  n_gpus = 2, 4 or 8 (world_size must
     then be 0 or 1): the state's top log2(n_gpus) qubits select the device, the library runs one host thread per device
     and its own NCCL communicator / peer mappings inside, and every entry point below works on the WHOLE state (offsets and
     counts are global, results are returned once).  This is the mode a JVM host uses (one `submit-circuit` caller, one job
     id: application/backend.clj:72-112).  n_gpus = 0 or 1: one device (or, with world_size > 1, one SPMD rank per process as
     launched by torchrun). */
  int32_t n_gpus;
  int32_t device_ids[8];   /* CUDA ordinals of the n_gpus devices, in slice order; all -1 (or n_gpus entries of -1) = 0 .. n_gpus-1 */
  int32_t reserved[3];
} qcb_config;

/* ---- gate vocabulary: every branch of apply-gate-to-state (domain/circuit.clj:964-1071) ---- */
enum qcb_op_kind {
  QCB_OP_I = 0,            /* :i   (strict_parity: rejected, circuit.clj:1072)                  */
  QCB_OP_X, QCB_OP_Y, QCB_OP_Z, QCB_OP_H,
  QCB_OP_S, QCB_OP_SDG, QCB_OP_T, QCB_OP_TDG,
  QCB_OP_RX, QCB_OP_RY, QCB_OP_RZ, QCB_OP_PHASE,       /* q[0]=target, angle                     */
  QCB_OP_CNOT, QCB_OP_CZ, QCB_OP_CY,                   /* q[0]=control, q[1]=target              */
  QCB_OP_CRX, QCB_OP_CRY, QCB_OP_CRZ,                  /* q[0]=control, q[1]=target, angle       */
  QCB_OP_SWAP, QCB_OP_ISWAP,                           /* q[0]=qubit1, q[1]=qubit2               */
  QCB_OP_TOFFOLI,                                      /* q[0]=control1, q[1]=control2, q[2]=target */
  QCB_OP_FREDKIN,                                      /* q[0]=control, q[1]=target1, q[2]=target2  */
  QCB_OP_RYDBERG_CZ,                                   /* = CZ (gate.clj:999-1001)               */
  QCB_OP_RYDBERG_CPHASE,                               /* q[0]=control, q[1]=target, angle       */
  QCB_OP_RYDBERG_BLOCKADE,                             /* mask = set of qubits (bit q = qubit q), angle */
  QCB_OP_GLOBAL_H, QCB_OP_GLOBAL_X, QCB_OP_GLOBAL_Y, QCB_OP_GLOBAL_Z,
  QCB_OP_GLOBAL_RX, QCB_OP_GLOBAL_RY, QCB_OP_GLOBAL_RZ, /* angle                                 */
  /* generic forms (superset; used by the Kraus/noise path and by callers that pre-multiply) */
  QCB_OP_U1Q,              /* q[0]=target, mat = row-major 2x2 complex                          */
  QCB_OP_CU1Q,             /* q[0]=control, q[1]=target, mat applied as given (no transpose)    */
  QCB_OP_U2Q,              /* q[0],q[1] = targets (q[0] = more significant of the 4x4 basis), ext -> 32 doubles row-major */
  QCB_OP_MCPHASE,          /* mask = qubit set: multiply by e^{i angle} where ALL listed qubits are 1 (multi-controlled Z/phase) */
  QCB_OP_PHASE_ORACLE,     /* mask = basis-state index to mark: amplitude[index] *= -1 (Grover oracle, application/algorithm/grover.clj:38-120 as an operator) */
  QCB_OP_GROVER_DIFFUSION, /* 2|s><s| - I over all qubits (grover.clj:122-190 as an operator)    */
  QCB_OP_MEASURE,          /* :measure (circuit.clj:1086-1092): ext -> int32[n_mask] qubit list (outcome bit i <->
                              i-th listed qubit), angle = the uniform draw u in [0,1) (ignored by qcb_run_noisy,
                              which takes the draw from its stream); collapses + renormalises            */
  QCB_OP_KIND_COUNT
};

typedef struct qcb_op {
  int32_t  kind;       /* enum qcb_op_kind                                              */
  int32_t  q[3];       /* qubit operands (reference numbering), unused = -1             */
  int32_t  n_mask;     /* number of qubits in `mask` (blockade / mcphase), else 0       */
  int32_t  _pad;
  uint64_t mask;       /* qubit set or basis index, see kinds                            */
  double   angle;
  double   mat[8];     /* 2x2 complex row-major for U1Q / CU1Q                           */
  const void* ext;     /* U2Q: 32 doubles; MEASURE: int32[n_mask]; otherwise NULL. Only read during the call. */
} qcb_op;

/* ---- library-level ---- */
int32_t qcb_abi_version(void);
/* message of the last failing call on this handle (or the last failing qcb_create when h == NULL) */
int32_t qcb_last_error(qcb_handle h, char* buf, size_t len);
int32_t qcb_device_count(int32_t* count);
int32_t qcb_nccl_unique_id(void* out128);            /* rank 0 creates, plumbing broadcasts   */

/* ---- lifecycle ---- */
int32_t qcb_config_default(qcb_config* cfg);
int32_t qcb_create(const qcb_config* cfg, qcb_handle* out);
int32_t qcb_destroy(qcb_handle h);
int32_t qcb_synchronize(qcb_handle h);

/* ---- state: zero-state / computational-basis-state / :initial-state (domain/state.clj:251-284, 484-519;
        ideal_simulator.clj:85) ---- */
int32_t qcb_set_zero(qcb_handle h);
int32_t qcb_set_basis(qcb_handle h, uint64_t index);
/* full (or, multi-GPU, this rank's slice [rank*2^(n-p), ...)) state from host memory */
int32_t qcb_set_state(qcb_handle h, const double* host_amps, uint64_t count);
int32_t qcb_get_state(qcb_handle h, uint64_t offset, uint64_t count, double* out_amps);
int32_t qcb_get_amplitudes(qcb_handle h, const uint64_t* indices, uint64_t n, double* out_amps); /* result.clj:393-402 */
int32_t qcb_normalize(qcb_handle h);                 /* normalize-state, domain/state.clj:544-551 */
/* raw device pointer of the local slice (for zero-copy plumbing: torch views, peer access); a multi-GPU handle (n_gpus > 1)
   returns the slice of device `device_ids[0]` */
int32_t qcb_state_dev_ptr(qcb_handle h, void** dev_ptr, uint64_t* local_count);

/* ---- gates: replaces (reduce apply-operation-to-state state ops), domain/circuit.clj:1782 ---- */
int32_t qcb_apply_ops(qcb_handle h, const qcb_op* ops, uint64_t n_ops);

/* ---- measurement: domain/state.clj:651-682, 894-913, 946-1014; domain/result.clj:201-252 ---- */
int32_t qcb_norm2(qcb_handle h, double* out_norm);                       /* sqrt(sum |a|^2)          */
int32_t qcb_probabilities(qcb_handle h, uint64_t offset, uint64_t count, double* out_probs);
/* measure-state rule on caller-supplied uniforms u in [0,1): outcome = #{i : cum_i < total*u} clamped */
int32_t qcb_sample(qcb_handle h, const double* uniforms, uint64_t n_shots, uint64_t* outcomes);
/* measure-specific-qubits: marginal over `qubits` (outcome bit i <-> qubits[i]), one draw, collapse + renormalise */
int32_t qcb_measure_qubits(qcb_handle h, const int32_t* qubits, int32_t m, double u, int32_t* out_bits, double* out_prob);
/* marginal distribution only (no collapse): out_probs[2^m] */
int32_t qcb_marginal_probabilities(qcb_handle h, const int32_t* qubits, int32_t m, double* out_probs);

/* ---- expectation: domain/observables.clj:216-251, domain/hamiltonian.clj:91-114, result.clj:266-288 ----
   Sharded states: a Pauli string with X / Y factors on global qubits is evaluated after the term's qubits have been brought
   into local positions by qubit exchanges (the layout permutation is tracked; nothing is moved back until a read needs the
   canonical order).  Only a string with more X / Y factors than one GPU holds qubits fails (QCB_ERR_UNSUPPORTED). */
int32_t qcb_expect_pauli(qcb_handle h, const char* pauli_string, double* out);
int32_t qcb_expect_hamiltonian(qcb_handle h, const double* coeffs, const char* const* pauli_strings,
                               uint64_t n_terms, double* out_energy, double* out_terms /* may be NULL */);
int32_t qcb_expect_1q(qcb_handle h, const double mat[8], int32_t target, double* out);
/* |<psi|phi>| against a host reference state (state-fidelity, domain/state.clj:1176-1185) */
int32_t qcb_fidelity(qcb_handle h, const double* host_amps, uint64_t count, double* out);

/* ---- noise: domain/channel.clj:162-244, domain/noise.clj:65-202,
        adapter/backend/hardware_simulator.clj:84-191 ---- */
int32_t qcb_apply_kraus_1q(qcb_handle h, const double mat[8], int32_t target);  /* K psi / ||K psi|| */

/* One noise entry per gate kind (the reference looks noise up by :operation-type, noise.clj:69-71). */
typedef struct qcb_noise_entry {
  int32_t op_kind;        /* enum qcb_op_kind this entry applies to                                    */
  int32_t n_kraus;        /* 1..4 Kraus operators                                                      */
  double  kraus[4][8];    /* row-major 2x2 complex each, coefficients included (channel.clj:52-121)    */
} qcb_noise_entry;

typedef struct qcb_noise_table {
  const qcb_noise_entry* entries;
  int32_t n_entries;
  int32_t has_readout;            /* :readout-error present                                            */
  double  prob_0_to_1, prob_1_to_0;
  const double* correlation;      /* n x n row-major factors [src*n + dst] or NULL (noise.clj:136-145) */
} qcb_noise_table;

/*
 * Per-shot trajectory loop of execute-circuit-simulation-with-trajectories (hardware_simulator.clj:120-185).
 * uniforms: [n_shots x draws_per_shot] row-major, consumed per shot in the reference's order: circuit
 * order (one draw per multi-Kraus noisy gate), one for the final measurement, n_qubits for readout when
 * has_readout.  out_outcomes[shot] = measured index AFTER readout flips (MSB-first bitstring as integer).
 * trajectories_out (may be NULL): first min(n_shots, max_traj) final states, 2^n complex each.
 * The handle's state holds the last shot's final state afterwards (:final-state).
 */
/* :initial-state of the trajectory loop ((or (:initial-state options) zero-state), hardware_simulator.clj:128-131): every
   trajectory of the following qcb_run_noisy calls starts from these amplitudes (2^n complex, kept on the device) instead
   of |0...0>; host_amps == NULL restores |0...0>. */
int32_t qcb_noisy_set_initial_state(qcb_handle h, const double* host_amps, uint64_t count);
int32_t qcb_noisy_draws_per_shot(qcb_handle h, const qcb_op* ops, uint64_t n_ops, const qcb_noise_table* noise,
                                 uint64_t* out_draws);
int32_t qcb_run_noisy(qcb_handle h, const qcb_op* ops, uint64_t n_ops, const qcb_noise_table* noise,
                      const double* uniforms, uint64_t draws_per_shot, uint64_t n_shots,
                      uint64_t* out_outcomes, double* trajectories_out, uint64_t max_traj);

/* ---- execution statistics of the last qcb_apply_ops / qcb_run_noisy on this handle ---- */
typedef struct qcb_stats {
  uint64_t n_ops;             /* ops submitted                                                  */
  uint64_t n_gates_lowered;   /* after expanding global-* gates                                 */
  uint64_t n_sweeps;          /* fused tile sweeps (kernel launches of the gate executor)       */
  uint64_t n_rounds;          /* dense-block rounds inside those sweeps (a paired pass counts two)  */
  uint64_t n_kernel_launches; /* every kernel this library launched for the call                */
  uint64_t n_exchanges;       /* global<->local qubit swaps (multi-GPU)                         */
  uint64_t bytes_exchanged;   /* bytes this rank sent over NVLink                               */
  double   algorithmic_bytes; /* sum over sweeps of bytes they must move (32 * 2^n_local each)  */
  double   unfused_bytes;     /* sum over ORIGINAL gates of 32*2^n*f(g)            */
  double   gpu_ms;            /* device time of the call (CUDA events on the handle's stream)   */
  double   exchange_ms;       /* device time spent in exchanges                                 */
} qcb_stats;
int32_t qcb_get_stats(qcb_handle h, qcb_stats* out);
/* CUDA-event stopwatch on the handle's own stream (the stream every kernel of this handle is launched on):
   start records an event; stop records a second one, synchronises on it and returns the elapsed device ms. */
int32_t qcb_timer_start(qcb_handle h);
int32_t qcb_timer_stop(qcb_handle h, double* out_ms);

/* ---- host-only planning API (no GPU needed): what the scheduler would do with an op list.
        Used by the CPU test-suite and by INTEGRATION diagnostics. ---- */
typedef struct qcb_plan qcb_plan;
int32_t qcb_plan_create(const qcb_config* cfg, const qcb_op* ops, uint64_t n_ops, qcb_plan** out);
/* plans `ops` by REPLAYING the scheduler decisions recorded on `ops_recorded` (same gates and qubits, other angles):
   what a handle does when a variational loop (VQE / QAOA objective, application/algorithm/variational_algorithm.clj:
   330-360) submits the same ansatz again.  Fails with QCB_ERR_INVALID when the two lists differ in structure. */
int32_t qcb_plan_create_replayed(const qcb_config* cfg, const qcb_op* ops_recorded, const qcb_op* ops, uint64_t n_ops, qcb_plan** out);
int32_t qcb_plan_destroy(qcb_plan* p);
/* serialises the plan as a flat little-endian word stream (layout documented in csrc/plan.h) */
int32_t qcb_plan_serialize(const qcb_plan* p, uint64_t* out_words, uint64_t capacity, uint64_t* n_words);
int32_t qcb_plan_summary(const qcb_plan* p, uint64_t* n_stages, uint64_t* n_rounds, uint64_t* n_exchanges);

/* ---- jobs: LocalQuantumSimulator (adapter/backend/ideal_simulator.clj:100-176) ---- */
#define QCB_JOB_QUEUED 0
#define QCB_JOB_RUNNING 1
#define QCB_JOB_COMPLETED 2
#define QCB_JOB_FAILED 3
#define QCB_JOB_CANCELLED 4
#define QCB_JOB_NOT_FOUND 5

typedef struct qcb_job_request {
  const qcb_op* ops; uint64_t n_ops;          /* copied at submit                                   */
  const double* initial_state; uint64_t initial_count; /* NULL = |0...0>                            */
  const double* uniforms; uint64_t n_shots;   /* measurement shots (copied); 0 = none               */
  const double* ham_coeffs; const char* const* ham_strings; uint64_t n_terms; /* optional energy    */
  int32_t want_probabilities;                 /* keep |a|^2 (2^n doubles) in the result             */
  int32_t want_state;                         /* keep the final state (2^n complex) in the result   */
} qcb_job_request;

typedef struct qcb_job_result {
  int32_t status;
  double  execution_time_ms;
  uint64_t n_shots; uint64_t* outcomes;       /* caller-allocated [n_shots]                          */
  double  energy; int32_t has_energy;
  double* probabilities; uint64_t prob_capacity;   /* caller-allocated, may be NULL                  */
  double* state; uint64_t state_capacity;          /* caller-allocated (complex count), may be NULL  */
  char    error_message[256];
} qcb_job_result;

int32_t qcb_submit(qcb_handle h, const qcb_job_request* req, uint64_t* out_job_id);
int32_t qcb_job_status(qcb_handle h, uint64_t job_id, int32_t* out_status);
int32_t qcb_job_result_get(qcb_handle h, uint64_t job_id, qcb_job_result* inout);
int32_t qcb_cancel(qcb_handle h, uint64_t job_id, int32_t* out_status);
/* Drops a finished job and everything it holds (outcomes, probabilities, final state); later queries answer
   QCB_JOB_NOT_FOUND.  Without it the library keeps the payload of the 64 most recently finished jobs only (older ones keep
   their status and timing, their buffers are freed) - a VQE / QAOA loop through the job API does not grow the host heap. */
int32_t qcb_job_release(qcb_handle h, uint64_t job_id);
int32_t qcb_queue_status(qcb_handle h, uint64_t* queued, uint64_t* running, uint64_t* completed);

/* ---- P2: small dense complex linear algebra (domain/math/protocols.clj MatrixAlgebra subset used on the
        simulation path: matrix-vector-product, matrix-multiply, kronecker-product, inner-product,
        outer-product, trace, norm2, add, scale).  Row-major interleaved arrays, computed on the GPU. ---- */
int32_t qcb_la_matvec(qcb_handle h, const double* A, const double* x, uint64_t rows, uint64_t cols, double* y);
int32_t qcb_la_matmul(qcb_handle h, const double* A, const double* B, uint64_t m, uint64_t k, uint64_t n, double* C);
int32_t qcb_la_kron(qcb_handle h, const double* A, uint64_t ar, uint64_t ac, const double* B, uint64_t br, uint64_t bc, double* C);
int32_t qcb_la_inner(qcb_handle h, const double* x, const double* y, uint64_t n, double out[2]);   /* conj(x) . y */
int32_t qcb_la_outer(qcb_handle h, const double* x, const double* y, uint64_t n, uint64_t m, double* C); /* x y^H */
int32_t qcb_la_trace(qcb_handle h, const double* A, uint64_t n, double out[2]);
int32_t qcb_la_norm2(qcb_handle h, const double* x, uint64_t n, double* out);
int32_t qcb_la_axpby(qcb_handle h, const double alpha[2], const double* x, const double beta[2], const double* y, uint64_t n, double* out);

/* ---- P2, remaining protocol methods (domain/math/protocols.clj:81-521; reference implementation
        domain/math/fastmath/complex_linear_algebra.clj:174-1470).  Small dense matrices: computed on the HOST inside the
        library (cyclic Jacobi, Householder, shifted QR), no GPU work, `h` may be NULL.  Row-major interleaved buffers,
        caller-allocated outputs.  Conventions: eigenvalues ascending (general: by real part, then imaginary part),
        eigenvector k stored at [k*n, (k+1)*n) and normalised, singular values descending with FULL U (m x m) and
        V^H (n x n), A = P L U, A = Q R with Q m x m, A = L L^H, principal branches for log and sqrt. ---- */
int32_t qcb_la_hadamard(qcb_handle h, const double* A, const double* B, uint64_t n_elements, double* C);          /* hadamard-product :213 */
int32_t qcb_la_transpose(qcb_handle h, const double* A, uint64_t rows, uint64_t cols, int32_t conjugate, double* out); /* transpose :239, conjugate-transpose :249 */
int32_t qcb_la_solve(qcb_handle h, const double* A, const double* B, uint64_t n, uint64_t nrhs, double* X);        /* solve-linear-system :284 */
int32_t qcb_la_inverse(qcb_handle h, const double* A, uint64_t n, double* out);                                    /* inverse :295 */
int32_t qcb_la_is_hermitian(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out);                  /* hermitian? :306 */
int32_t qcb_la_is_diagonal(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out);                   /* diagonal? :317 */
int32_t qcb_la_is_unitary(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out);                    /* unitary? :328 */
int32_t qcb_la_is_positive_semidefinite(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out);      /* positive-semidefinite? :339 (error if not Hermitian) */
int32_t qcb_la_eigen_hermitian(qcb_handle h, const double* A, uint64_t n, double* eigenvalues, double* eigenvectors); /* eigen-hermitian :363 */
int32_t qcb_la_eigen_general(qcb_handle h, const double* A, uint64_t n, double* eigenvalues /* complex[n] */, double* eigenvectors); /* eigen-general :375 */
int32_t qcb_la_svd(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* U, double* S, double* Vh);       /* svd :390 */
int32_t qcb_la_lu(qcb_handle h, const double* A, uint64_t n, double* P, double* L, double* U);                     /* lu-decomposition :403 */
int32_t qcb_la_qr(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* Q, double* R);                    /* qr-decomposition :416 */
int32_t qcb_la_cholesky(qcb_handle h, const double* A, uint64_t n, double* L);                                     /* cholesky-decomposition :431 */
int32_t qcb_la_matrix_exp(qcb_handle h, const double* A, uint64_t n, double* out);                                 /* matrix-exp :453 */
int32_t qcb_la_matrix_log(qcb_handle h, const double* A, uint64_t n, double* out);                                 /* matrix-log :466 */
int32_t qcb_la_matrix_sqrt(qcb_handle h, const double* A, uint64_t n, double* out);                                /* matrix-sqrt :479 */
int32_t qcb_la_spectral_norm(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* out);                  /* spectral-norm :500 */
int32_t qcb_la_condition_number(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* out);               /* condition-number :510 */

#ifdef __cplusplus
}
#endif
#endif /* QCB200_H */
