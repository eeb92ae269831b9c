// sim.cu — the handle behind include/qcb200.h: device state, plan execution, measurement, expectation,
// noise trajectories, qubit exchanges (NCCL) and the job layer.  No CPU fallback anywhere: every entry
// point that touches the state needs the CUDA device the handle was created on.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/qcb200.h"
#include "kernels.h"
#include "la_host.h"
#include "plan.h"

using namespace qcb;

// ------------------------------------------------------------------ NCCL, resolved at run time
namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool load_nccl(std::string& err) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // reuse the copy torch already loaded
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
  g_nccl.lib = lib;
#define QCB_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));       \
  if (!g_nccl.field) { err = std::string("missing NCCL symbol ") + name; return false; }
  QCB_SYM(GetUniqueId, "ncclGetUniqueId");
  QCB_SYM(CommInitRank, "ncclCommInitRank");
  QCB_SYM(CommDestroy, "ncclCommDestroy");
  QCB_SYM(Send, "ncclSend");
  QCB_SYM(Recv, "ncclRecv");
  QCB_SYM(GroupStart, "ncclGroupStart");
  QCB_SYM(GroupEnd, "ncclGroupEnd");
  QCB_SYM(AllReduce, "ncclAllReduce");
  QCB_SYM(AllGather, "ncclAllGather");
  QCB_SYM(GetErrorString, "ncclGetErrorString");
#undef QCB_SYM
  g_nccl.ok = true;
  return true;
}

std::string g_create_error;
std::mutex g_create_mu;
}  // namespace

// ------------------------------------------------------------------ jobs
struct Job {
  uint64_t id = 0;
  std::atomic<int> status{QCB_JOB_QUEUED};
  std::atomic<bool> cancel{false};
  std::vector<qcb_op> ops;
  std::vector<std::vector<double>> ext_d;     // owned copies of op.ext payloads
  std::vector<std::vector<int32_t>> ext_i;
  std::vector<double> initial;
  std::vector<double> uniforms;
  std::vector<double> ham_coeffs;
  std::vector<std::string> ham_strings;
  int want_probs = 0, want_state = 0;
  // results
  double exec_ms = 0;
  std::vector<uint64_t> outcomes;
  double energy = 0; int has_energy = 0;
  std::vector<double> probs, state;
  bool payload_dropped = false;               // pruned: outcomes / probabilities / state are gone (64 newest finished jobs keep theirs)
  std::string error;
};

// ------------------------------------------------------------------ single-process multi-GPU: what the member handles share
// One qcb_handle created with qcb_config.n_gpus > 1 is a *group*: it owns one ordinary rank handle per device (the same
// SPMD code path torchrun ranks run, one host thread per device for every call) and presents the whole state.
struct GroupShared {
  int n = 0;
  std::vector<double2*> states;                   // state allocation of every member (direct peer access inside one process)
  std::vector<int> devices;
  std::mutex mu; std::condition_variable cv;
  int arrived = 0; uint64_t gen = 0; bool all_ok = true, verdict = true;
  // host barrier that also agrees on a flag: returns true when EVERY member passed ok = true
  bool agree(bool ok) {
    std::unique_lock<std::mutex> lk(mu);
    all_ok = all_ok && ok;
    const uint64_t my_gen = gen;
    if (++arrived == n) { verdict = all_ok; all_ok = true; arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != my_gen; });
    return verdict;
  }
};

// ------------------------------------------------------------------ the handle
struct qcb_sim {
  // logical bits that can be 1 where the state is non-zero (plan.h: Plan::support_in); ~0 = nothing known.  0 only straight after
  // qcb_set_zero with QCB_ZERO_SKIP=1 on a single-GPU handle; every API entry except qcb_apply_ops forgets it (ENTER).
  uint64_t support = ~0ULL;
  // group handle (n_gpus > 1): no device state of its own, every entry point fans out to `members`
  bool is_group = false;
  std::vector<qcb_sim*> members;
  std::shared_ptr<GroupShared> gs;                // set on group handles and on their members
  Config cfg;
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr;
  double2* state = nullptr;
  uint64_t local_count = 0;
  uint64_t* d_prog = nullptr; uint64_t* h_prog = nullptr; size_t prog_cap = 0;   // words
  cudaEvent_t prog_ev = nullptr; bool prog_ev_valid = false;
  double* d_vals = nullptr;                       // 256 doubles: device-side coefficients / small results
  TileMaps maps;                                  // TMA tensor maps of `state` (one per run length)
  double* d_partials = nullptr; size_t partials_cap = 0;   // doubles
  unsigned char* d_scratch = nullptr; size_t scratch_cap = 0;  // bytes
  unsigned char* h_pin = nullptr; size_t pin_cap = 0;          // bytes (pinned)
  std::vector<int> perm;                          // logical bit -> physical bit
  // plan traces by circuit structure (plan.h: PlanTrace): a variational loop re-plans the same ansatz with new angles
  // every evaluation; a hit replays the recorded decisions and only rebuilds the matrices
  std::unordered_map<uint64_t, std::vector<std::shared_ptr<PlanTrace>>> traces;
  size_t n_traces = 0, trace_words = 0;
  uint64_t trace_hits = 0, trace_misses = 0;
  qcb_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; bool timing_pending = false;
  cudaEvent_t xev0 = nullptr, xev1 = nullptr;
  cudaStream_t xstream = nullptr;                 // copy-back stream of the qubit exchange
  // peer-to-peer exchange: every rank's state allocation is IPC-mapped by its exchange partners (ranks differing in
  // one bit), so the moving half is pulled straight out of the partner's HBM over NVLink by the copy engines
  bool p2p = false;
  bool peers_ipc = false;                         // peer_state entries are IPC mappings (closed on destroy); else direct pointers
  int xmode = 0;                                  // 0 = in-place swap kernel over peer memory, 1 = copy-engine pull, 2 = NCCL send/recv
  std::vector<double2*> peer_state;               // [world], nullptr = not mapped
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> xtimes;   // exchange event pairs of the current call (resolved by qcb_get_stats)
  std::vector<cudaEvent_t> xev_pool;
  double2* noisy_init = nullptr;                  // qcb_noisy_set_initial_state: device copy of the trajectories' initial state
  std::vector<double2*> ckpts;                    // state checkpoints of the noisy trajectory tree (kept between calls)
  cudaEvent_t xrecv[2] = {nullptr, nullptr}, xcopy[2] = {nullptr, nullptr}, xpack[2] = {nullptr, nullptr};
  cudaEvent_t tev0 = nullptr, tev1 = nullptr;
  std::string err;
  std::recursive_mutex mu;
  // multi-GPU
  ncclComm_t comm = nullptr;
  double2* xbuf = nullptr; uint64_t xbuf_count = 0;
  // jobs
  std::mutex jmu; std::condition_variable jcv;
  std::map<uint64_t, std::shared_ptr<Job>> jobs;
  std::deque<std::shared_ptr<Job>> queue;
  std::deque<uint64_t> finished;                  // ids of finished jobs whose payload is still held (newest last)
  std::thread worker; bool worker_started = false; bool stopping = false;
  uint64_t next_job = 1;
  std::atomic<bool>* active_cancel = nullptr;
};

namespace {

int fail(qcb_sim* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else { std::lock_guard<std::mutex> lk(g_create_mu); g_create_error = msg; }
  return code;
}

#define CU(h, expr)                                                                                   \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return fail(h, QCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

#define NC(h, expr)                                                                                   \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess)                                                                            \
      return fail(h, QCB_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));            \
  } while (0)

#define RET(expr)                  \
  do {                             \
    int _rc = (expr);              \
    if (_rc != QCB_OK) return _rc; \
  } while (0)

int red_grid(const qcb_sim* h) { return h->num_sms * 4; }

int ensure_partials(qcb_sim* h, size_t doubles) {
  if (doubles <= h->partials_cap) return QCB_OK;
  if (h->d_partials) cudaFree(h->d_partials);
  h->d_partials = nullptr; h->partials_cap = 0;
  CU(h, cudaMalloc(&h->d_partials, doubles * sizeof(double)));
  h->partials_cap = doubles;
  return QCB_OK;
}
int ensure_scratch(qcb_sim* h, size_t bytes) {
  if (bytes <= h->scratch_cap) return QCB_OK;
  if (h->d_scratch) cudaFree(h->d_scratch);
  h->d_scratch = nullptr; h->scratch_cap = 0;
  CU(h, cudaMalloc(&h->d_scratch, bytes));
  h->scratch_cap = bytes;
  return QCB_OK;
}
int ensure_pinned(qcb_sim* h, size_t bytes) {
  if (bytes <= h->pin_cap) return QCB_OK;
  if (h->h_pin) cudaFreeHost(h->h_pin);
  h->h_pin = nullptr; h->pin_cap = 0;
  CU(h, cudaMallocHost(&h->h_pin, bytes));
  h->pin_cap = bytes;
  return QCB_OK;
}
int ensure_prog(qcb_sim* h, size_t words) {
  if (words <= h->prog_cap) return QCB_OK;
  size_t cap = std::max<size_t>(words, 1 << 16);
  if (h->d_prog) cudaFree(h->d_prog);
  if (h->h_prog) cudaFreeHost(h->h_prog);
  h->d_prog = nullptr; h->h_prog = nullptr; h->prog_cap = 0;
  CU(h, cudaMalloc(&h->d_prog, cap * sizeof(uint64_t)));
  CU(h, cudaMallocHost(&h->h_prog, cap * sizeof(uint64_t)));
  h->prog_cap = cap;
  return QCB_OK;
}

int do_exchange_nccl(qcb_sim* h, int gbit, int lbit);

// Exchange timing without host synchronisation: an event pair per exchange on the handle's stream, resolved when the
// statistics are read (qcb_get_stats), so that the scheduler keeps planning while an exchange is in flight.
cudaEvent_t xev_get(qcb_sim* h) {
  if (!h->xev_pool.empty()) { cudaEvent_t e = h->xev_pool.back(); h->xev_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
int xtime_begin(qcb_sim* h) {
  cudaEvent_t e = xev_get(h);
  if (!e) return fail(h, QCB_ERR_CUDA, "cudaEventCreate failed");
  CU(h, cudaEventRecord(e, h->stream));
  h->xtimes.emplace_back(e, nullptr);
  return QCB_OK;
}
int xtime_end(qcb_sim* h) {
  cudaEvent_t e = xev_get(h);
  if (!e) return fail(h, QCB_ERR_CUDA, "cudaEventCreate failed");
  CU(h, cudaEventRecord(e, h->stream));
  h->xtimes.back().second = e;
  return QCB_OK;
}
// adds the elapsed time of every finished exchange to stats.exchange_ms (synchronises on their end events)
void xtime_resolve(qcb_sim* h) {
  for (auto& pr : h->xtimes) {
    if (pr.first && pr.second && cudaEventSynchronize(pr.second) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) h->stats.exchange_ms += ms;
    }
    if (pr.first) h->xev_pool.push_back(pr.first);
    if (pr.second) h->xev_pool.push_back(pr.second);
  }
  h->xtimes.clear();
}

int ensure_exchange_buffers(qcb_sim* h, uint64_t half) {
  if (h->xbuf) return QCB_OK;
  static const int xlog = [] { const char* e = getenv("QCB_XCHUNK_LOG2"); int v = e ? atoi(e) : 25; return v < 4 ? 4 : (v > 28 ? 28 : v); }();
  uint64_t cnt = std::min<uint64_t>(half, 1ULL << xlog);        // 2 send + 2 receive buffers of <= 512 MiB
  CU(h, cudaMalloc(&h->xbuf, 4 * cnt * sizeof(double2)));
  h->xbuf_count = cnt;
  CU(h, cudaStreamCreateWithFlags(&h->xstream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CU(h, cudaEventCreateWithFlags(&h->xrecv[i], cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->xcopy[i], cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->xpack[i], cudaEventDisableTiming));
  }
  return QCB_OK;
}

// Peer-to-peer variant of the exchange: each rank PULLS the partner's moving half out of the partner's HBM
// (IPC-mapped at create time) into a local staging buffer with the copy engines, then scatters it into its own vacated
// positions.  Cross-process ordering rides on one-double ncclSend/ncclRecv handshakes, which are stream-ordered on both
// sides: H0 = "my earlier kernels are done, you may read my state", HS(c) = "I have pulled chunk c of yours, you may
// overwrite it".  A plain peer copy reaches ~770 GB/s per direction on this pool; ncclSend/ncclRecv ~520.
int do_exchange_p2p(qcb_sim* h, int gbit, int lbit) {
  const int nl = h->cfg.n_local;
  const int j = gbit - nl;
  const int myb = (h->cfg.rank >> j) & 1;
  const int peer = h->cfg.rank ^ (1 << j);
  const uint64_t want = (uint64_t)(1 - myb), want_p = (uint64_t)myb;   // moving half: mine has bit lbit == !myb, the partner's == myb
  const uint64_t half = h->local_count >> 1, block = 1ULL << lbit;
  RET(ensure_exchange_buffers(h, half));
  const double2* pstate = h->peer_state[peer];
  const uint64_t chunk = h->xbuf_count;
  const uint64_t n_chunks = (half + chunk - 1) / chunk;
  const bool contiguous = block >= chunk;
  auto pos = [&](const double2* base, uint64_t sel, uint64_t first) {      // address of moving-half offset `first`
    const uint64_t bi = first >> lbit, in = first & (block - 1);
    return base + (((bi << 1) | sel) << lbit) + in;
  };
  auto recvbuf = [&](uint64_t c) { return h->xbuf + (2 + (c & 1)) * chunk; };
  auto count_of = [&](uint64_t c) { return std::min<uint64_t>(chunk, half - c * chunk); };
  double* token = h->d_vals + 200;
  auto handshake = [&]() -> int {
    NC(h, g_nccl.GroupStart());
    NC(h, g_nccl.Send(token, 1, ncclDouble, peer, h->comm, h->stream));
    NC(h, g_nccl.Recv(token + 1, 1, ncclDouble, peer, h->comm, h->stream));
    NC(h, g_nccl.GroupEnd());
    return QCB_OK;
  };
  RET(xtime_begin(h));
  RET(handshake());                                                         // H0
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const int sb = (int)(c & 1);
    const uint64_t cnt = count_of(c), first = c * chunk;
    if (c >= 2) CU(h, cudaStreamWaitEvent(h->stream, h->xcopy[sb], 0));     // staging buffer free again
    if (contiguous)
      CU(h, cudaMemcpyAsync(recvbuf(c), pos(pstate, want_p, first), cnt * sizeof(double2), cudaMemcpyDefault, h->stream));
    else
      CU(h, cudaMemcpy2DAsync(recvbuf(c), block * sizeof(double2), pos(pstate, want_p, first), 2 * block * sizeof(double2),
                              block * sizeof(double2), cnt >> lbit, cudaMemcpyDefault, h->stream));
    RET(handshake());                                                       // HS(c)
    CU(h, cudaEventRecord(h->xrecv[sb], h->stream));
    CU(h, cudaStreamWaitEvent(h->xstream, h->xrecv[sb], 0));
    double2* dst = h->state + (pos(h->state, want, first) - h->state);
    if (contiguous)
      CU(h, cudaMemcpyAsync(dst, recvbuf(c), cnt * sizeof(double2), cudaMemcpyDeviceToDevice, h->xstream));
    else
      CU(h, cudaMemcpy2DAsync(dst, 2 * block * sizeof(double2), recvbuf(c), block * sizeof(double2), block * sizeof(double2),
                              cnt >> lbit, cudaMemcpyDeviceToDevice, h->xstream));
    CU(h, cudaEventRecord(h->xcopy[sb], h->xstream));
    h->stats.bytes_exchanged += cnt * sizeof(double2);
  }
  for (uint64_t i = 0; i < 2 && i < n_chunks; ++i) CU(h, cudaStreamWaitEvent(h->stream, h->xcopy[i], 0));
  RET(xtime_end(h));
  h->stats.n_exchanges++;
  return QCB_OK;
}

// stream-ordered barrier across all ranks: a one-double all-reduce on the handle's stream completes only when every rank's
// stream has reached it
int stream_barrier(qcb_sim* h) {
  double* token = h->d_vals + 204;
  NC(h, g_nccl.AllReduce(token, token, 1, ncclDouble, ncclMax, h->comm, h->stream));
  return QCB_OK;
}

// Simultaneous swap of up to MAX_SWAP_BITS (global bit, local bit) pairs with disjoint bits: one in-place pass of
// k_swap_global over the peer-mapped slices, bracketed by stream-ordered barriers (nobody touches a peer's slice before
// the peer's earlier kernels are done; nobody reads its own slice before the peers' stores have landed).  No staging
// buffers, no host synchronisation.
int do_exchange_swap(qcb_sim* h, const std::pair<int, int>* pairs, int k) {
  const int nl = h->cfg.n_local;
  std::vector<int> gb(k), order(k);
  for (int j = 0; j < k; ++j) gb[j] = pairs[j].first;
  std::sort(gb.begin(), gb.end());                                            // group-value bit i <-> gb[i]
  for (int j = 0; j < k; ++j) order[j] = j;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return pairs[a].second < pairs[b].second; });
  SwapBits sb; sb.k = k;
  for (int j = 0; j < MAX_SWAP_BITS; ++j) { sb.lpos[j] = 0; sb.pair[j] = 0; }
  for (int j = 0; j < k; ++j) {
    sb.lpos[j] = pairs[order[j]].second;
    sb.pair[j] = (int)(std::find(gb.begin(), gb.end(), pairs[order[j]].first) - gb.begin());
  }
  uint32_t g = 0;
  for (int i = 0; i < k; ++i) g |= (uint32_t)((h->cfg.rank >> (gb[i] - nl)) & 1) << i;
  SwapPeers peers;
  for (int v = 0; v < (1 << MAX_SWAP_BITS); ++v) peers.p[v] = nullptr;
  for (uint32_t v = 0; v < (1u << k); ++v) {
    if (v == g) continue;
    int r = h->cfg.rank;
    for (int i = 0; i < k; ++i) r = (r & ~(1 << (gb[i] - nl))) | (int)((v >> i) & 1u) << (gb[i] - nl);
    if (!h->peer_state[r]) return fail(h, QCB_ERR_NCCL, "exchange partner " + std::to_string(r) + " is not peer-mapped");
    peers.p[v] = h->peer_state[r];
  }
  const uint64_t n_rest_half = h->local_count >> (k + 1);
  RET(stream_barrier(h));
  RET(xtime_begin(h));
  const uint64_t total = n_rest_half * ((1ull << k) - 1ull);
  const int grid = (int)std::min<uint64_t>((total + RED_THREADS - 1) / RED_THREADS, (uint64_t)h->num_sms * 8);
  if (grid > 0) CU(h, launch_swap_global(h->state, peers, sb, g, n_rest_half, grid, h->stream));
  RET(stream_barrier(h));
  RET(xtime_end(h));
  h->stats.n_kernel_launches++;
  h->stats.n_exchanges += (uint64_t)k;
  h->stats.bytes_exchanged += (h->local_count - (h->local_count >> k)) * sizeof(double2);
  return QCB_OK;
}

int do_exchange(qcb_sim* h, int gbit, int lbit) {
  if (h->p2p && h->xmode == 0) { const std::pair<int, int> pr(gbit, lbit); return do_exchange_swap(h, &pr, 1); }
  return (h->p2p && h->xmode == 1) ? do_exchange_p2p(h, gbit, lbit) : do_exchange_nccl(h, gbit, lbit);
}

// A run of exchanges with pairwise disjoint bits (what the scheduler emits "in one breath", and what restore_layout needs):
// with the swap kernel up to MAX_SWAP_BITS pairs move in one pass ((2^k - 1) / 2^k of a slice instead of k / 2)
int do_exchange_multi(qcb_sim* h, const std::vector<std::pair<int, int>>& pairs) {
  if (pairs.empty()) return QCB_OK;
  if (!(h->p2p && h->xmode == 0)) {
    for (auto& pr : pairs) RET(do_exchange(h, pr.first, pr.second));
    return QCB_OK;
  }
  static const int max_k = [] { const char* e = getenv("QCB_SWAP_BITS"); int v = e ? atoi(e) : MAX_SWAP_BITS; return v < 1 ? 1 : (v > MAX_SWAP_BITS ? MAX_SWAP_BITS : v); }();
  for (size_t i = 0; i < pairs.size(); i += (size_t)max_k)
    RET(do_exchange_swap(h, pairs.data() + i, (int)std::min<size_t>((size_t)max_k, pairs.size() - i)));
  return QCB_OK;
}

bool perm_is_identity(const qcb_sim* h) {
  for (size_t b = 0; b < h->perm.size(); ++b) if (h->perm[b] != (int)b) return false;
  return true;
}

// device -> host small copy through pinned memory, synchronising the stream
int read_back(qcb_sim* h, const void* dev, size_t bytes, void* host) {
  RET(ensure_pinned(h, bytes));
  CU(h, cudaMemcpyAsync(h->h_pin, dev, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  std::memcpy(host, h->h_pin, bytes);
  return QCB_OK;
}

// sum over ranks of `count` doubles at dev (in place)
int allreduce_sum(qcb_sim* h, double* dev, size_t count) {
  if (h->cfg.world <= 1) return QCB_OK;
  NC(h, g_nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, h->comm, h->stream));
  return QCB_OK;
}

// ---- exchange: swap global physical bit gbit with local physical bit lbit
int do_exchange_nccl(qcb_sim* h, int gbit, int lbit) {
  // Swap global physical bit gbit with local physical bit lbit: rank r and partner r ^ 2^j exchange the half of their
  // slice whose bit lbit differs from their own rank bit.  The half moves in chunks of <= 512 MiB through a two-deep
  // software pipeline: chunk c+1 is gathered into a send buffer (k_pack_half; skipped when the chunk is contiguous in
  // the state) and chunk c-1 is scattered back from its receive buffer (k_unpack_half / memcpy) on a second stream
  // while chunk c is on the wire as ONE ncclSend/ncclRecv pair (NCCL moves few large messages much faster than many
  // small ones: profiles/r1d_exchange_bench.log).
  const int nl = h->cfg.n_local;
  const int j = gbit - nl;
  const int myb = (h->cfg.rank >> j) & 1;
  const int peer = h->cfg.rank ^ (1 << j);
  const int want = 1 - myb;                         // the half of MY slice that moves: bit lbit == !myb
  const uint64_t half = h->local_count >> 1, block = 1ULL << lbit;
  RET(ensure_exchange_buffers(h, half));
  const uint64_t chunk = h->xbuf_count;              // amplitudes of the moving half per chunk
  const uint64_t n_chunks = (half + chunk - 1) / chunk;
  const bool contiguous = block >= chunk;            // a chunk lies inside one contiguous block of the state
  auto state_pos = [&](uint64_t first) {             // address of moving-half offset `first` (contiguous case)
    const uint64_t bi = first >> lbit, in = first & (block - 1);
    return h->state + (((bi << 1) | (uint64_t)want) << lbit) + in;
  };
  auto sendbuf = [&](uint64_t c) { return h->xbuf + (c & 1) * chunk; };
  auto recvbuf = [&](uint64_t c) { return h->xbuf + (2 + (c & 1)) * chunk; };
  auto count_of = [&](uint64_t c) { return std::min<uint64_t>(chunk, half - c * chunk); };
  RET(xtime_begin(h));
  CU(h, cudaEventRecord(h->xev0, h->stream));
  CU(h, cudaStreamWaitEvent(h->xstream, h->xev0, 0));                       // the state is final before anything is gathered
  auto pack = [&](uint64_t c) -> int {
    if (contiguous) return QCB_OK;
    CU(h, launch_pack_half(h->state, sendbuf(c), c * chunk, count_of(c), lbit, want, red_grid(h), h->xstream));
    CU(h, cudaEventRecord(h->xpack[c & 1], h->xstream));
    h->stats.n_kernel_launches++;
    return QCB_OK;
  };
  RET(pack(0));
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const int sb = (int)(c & 1);
    const uint64_t cnt = count_of(c);
    if (!contiguous) CU(h, cudaStreamWaitEvent(h->stream, h->xpack[sb], 0));
    if (c >= 2) CU(h, cudaStreamWaitEvent(h->stream, h->xcopy[sb], 0));       // receive buffer free again
    NC(h, g_nccl.GroupStart());
    NC(h, g_nccl.Send(contiguous ? state_pos(c * chunk) : sendbuf(c), cnt * 2, ncclDouble, peer, h->comm, h->stream));
    NC(h, g_nccl.Recv(recvbuf(c), cnt * 2, ncclDouble, peer, h->comm, h->stream));
    NC(h, g_nccl.GroupEnd());
    CU(h, cudaEventRecord(h->xrecv[sb], h->stream));
    if (c + 1 < n_chunks) RET(pack(c + 1));          // queued before the scatter of chunk c: overlaps the wire time of chunk c
    // partner's moving half (its bit lbit == myb) lands in MY moving positions (bit lbit == !myb), vacated by the send
    CU(h, cudaStreamWaitEvent(h->xstream, h->xrecv[sb], 0));
    if (contiguous) CU(h, cudaMemcpyAsync(state_pos(c * chunk), recvbuf(c), cnt * sizeof(double2), cudaMemcpyDeviceToDevice, h->xstream));
    else { CU(h, launch_unpack_half(h->state, recvbuf(c), c * chunk, cnt, lbit, want, red_grid(h), h->xstream)); h->stats.n_kernel_launches++; }
    CU(h, cudaEventRecord(h->xcopy[sb], h->xstream));
    h->stats.bytes_exchanged += cnt * sizeof(double2);
  }
  for (uint64_t i = 0; i < 2 && i < n_chunks; ++i) CU(h, cudaStreamWaitEvent(h->stream, h->xcopy[i], 0));
  RET(xtime_end(h));
  h->stats.n_exchanges++;
  return QCB_OK;
}

// ---- run a scheduled plan on the device
// Launches one planned stage.  Tile stages read their program from d_prog (uploaded by the caller at word offset `off`).
int execute_stage(qcb_sim* h, Plan& plan, size_t si) {
  if (h->active_cancel && h->active_cancel->load()) return fail(h, QCB_ERR_STATE, "cancelled");
  Stage& st = plan.stages[si];
  const uint64_t off = plan.stage_offsets[si];
  if (st.kind == S_TILE) {
    const uint32_t words = (uint32_t)plan.words[off + 2 + 42];      // descriptor part only (copied to smem)
    uint64_t active = 0;
    CU(h, launch_tile_stage(h->state, h->d_prog + off + 2, plan.words.data() + off + 2, words, h->d_vals, h->num_sms, h->stream, &active, &h->maps));
    if (active) { h->stats.n_sweeps++; h->stats.n_kernel_launches++; for (const Round& r : st.rounds) h->stats.n_rounds += r.dense_rounds(); }
  } else if (st.kind == S_SUM) {
    // Grover diffusion: sum of all amplitudes -> (alpha, beta) = (-1, 2*mean) in d_vals[0..4)
    const int grid = red_grid(h);
    RET(ensure_partials(h, (size_t)grid * 2));
    CU(h, launch_reduce(h->state, h->local_count, 0, h->d_partials, grid, h->stream));
    CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
    RET(allreduce_sum(h, h->d_vals + 8, 2));
    CU(h, launch_finalize(h->d_vals + 8, 1, 2, 1, std::ldexp(1.0, h->cfg.n_total), h->d_vals, h->stream));
    h->stats.n_kernel_launches += 3;
  } else if (st.kind == S_GROVER) {
    // diffusion + following phase oracles (+ the sum for the next diffusion) in one streaming pass
    const int grid = h->num_sms * 8;
    RET(ensure_partials(h, (size_t)grid * 2));
    GroverMarks mk; mk.n = (int)std::min<size_t>(st.marked.size(), 8);
    for (int k = 0; k < 8; ++k) mk.idx[k] = k < mk.n ? st.marked[k] : ~0ULL;
    CU(h, launch_grover_step(h->state, h->local_count, h->d_vals, mk, st.needs_sum ? h->d_partials : nullptr, grid, h->stream));
    h->stats.n_sweeps++; h->stats.n_kernel_launches++;
    if (st.needs_sum) {
      CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
      RET(allreduce_sum(h, h->d_vals + 8, 2));
      CU(h, launch_finalize(h->d_vals + 8, 1, 2, 1, std::ldexp(1.0, h->cfg.n_total), h->d_vals, h->stream));
      h->stats.n_kernel_launches += 2;
    }
  } else if (st.kind == S_EXCHANGE) {
    RET(do_exchange(h, st.gbit, st.lbit));         // (the streaming executor batches exchanges itself and never gets here)
  }
  return QCB_OK;
}

// Executes stages while the scheduler is still planning the rest of the circuit: every finished stage is copied into the
// pinned program buffer, uploaded and launched at once (kernel launches are asynchronous, so the host-side planning of
// stage k+1 overlaps the sweep of stage k).  The program buffers are sized for the whole op list up front: they must
// not move while earlier stages are in flight.
struct StreamingExecutor : StageSink {
  qcb_sim* h;
  std::vector<std::pair<int, int>> xq;             // consecutive exchange stages with pairwise disjoint bits, not yet executed
  explicit StreamingExecutor(qcb_sim* hh) : h(hh) {}
  int flush_exchanges() {
    if (xq.empty()) return QCB_OK;
    std::vector<std::pair<int, int>> q;
    q.swap(xq);
    return do_exchange_multi(h, q);
  }
  int on_stage(Plan& plan, size_t si) override {
    if (plan.stages[si].kind == S_EXCHANGE) {
      // consecutive exchanges on disjoint bits commute: queue them and move them in one pass
      const int gb = plan.stages[si].gbit, lb = plan.stages[si].lbit;
      for (auto& pr : xq) if (pr.first == gb || pr.second == lb || pr.first == lb || pr.second == gb) { RET(flush_exchanges()); break; }
      xq.emplace_back(gb, lb);
      if (h->active_cancel && h->active_cancel->load()) return fail(h, QCB_ERR_STATE, "cancelled");
      return QCB_OK;
    }
    RET(flush_exchanges());
    const size_t begin = plan.stage_offsets[si], end = plan.words.size();
    if (end > h->prog_cap) {
      // rare (bound below exceeded): let everything in flight finish, then grow
      CU(h, cudaStreamSynchronize(h->stream));
      RET(ensure_prog(h, end * 2));
      std::memcpy(h->h_prog, plan.words.data(), begin * sizeof(uint64_t));
    }
    std::memcpy(h->h_prog + begin, plan.words.data() + begin, (end - begin) * sizeof(uint64_t));
    CU(h, cudaMemcpyAsync(h->d_prog + begin, h->h_prog + begin, (end - begin) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    return execute_stage(h, plan, si);
  }
};

int run_gates(qcb_sim* h, std::vector<Gate>&& gates) {
  Plan plan;
  plan.cfg = h->cfg;
  plan.gates = std::move(gates);
  plan.support_in = (h->cfg.world == 1) ? h->support : ~0ULL;
  h->support = ~0ULL;                                   // until the plan has run to its end
  // program size bound: a tensor-core round holds at most 2^MAX_COND_BITS matrices of 256 doubles, and there are at most
  // as many rounds as gates; typical programs are ~80 words per gate
  RET(ensure_prog(h, std::max<size_t>(1 << 17, plan.gates.size() * 512)));
  if (h->prog_ev_valid) CU(h, cudaEventSynchronize(h->prog_ev));      // the previous call's uploads have left the pinned buffer
  StreamingExecutor sink(h);
  // plan-trace cache (QCB_PLAN_CACHE=0 disables it)
  static const bool cache_on = !(std::getenv("QCB_PLAN_CACHE") && std::atoi(std::getenv("QCB_PLAN_CACHE")) == 0);
  const bool cacheable = cache_on && plan.gates.size() >= 16 && plan.gates.size() <= (1u << 18);
  std::shared_ptr<PlanTrace> hit, rec;
  uint64_t hash = 1469598103934665603ULL;
  if (cacheable) {
    rec = std::make_shared<PlanTrace>();
    plan_structure_key(plan.cfg, plan.gates, h->perm, rec->key, plan.support_in);
    for (uint64_t w : rec->key) { hash ^= w; hash *= 1099511628211ULL; hash ^= hash >> 29; }
    auto it = h->traces.find(hash);
    if (it != h->traces.end())
      for (auto& t : it->second) if (t->key == rec->key) { hit = t; break; }
  }
  int rc = schedule(plan, h->perm, &sink, (cacheable && !hit) ? rec.get() : nullptr, hit.get());
  if (rc == QCB_OK) { rc = sink.flush_exchanges(); if (rc != QCB_OK) plan.error = "stage sink failed"; }
  CU(h, cudaEventRecord(h->prog_ev, h->stream));
  h->prog_ev_valid = true;
  if (rc != QCB_OK) {
    // a sink failure has already recorded its own message (CUDA / NCCL error); scheduler errors carry plan.error
    if (plan.error != "stage sink failed") return fail(h, rc, plan.error);
    return rc;
  }
  if (hit) ++h->trace_hits;
  else if (cacheable) {
    ++h->trace_misses;
    // bounded: at most 128 traces and 32 MB of keys per handle, then start over
    if (h->n_traces >= 128 || h->trace_words + rec->key.size() > (4u << 20)) { h->traces.clear(); h->n_traces = 0; h->trace_words = 0; }
    h->trace_words += rec->key.size();
    h->traces[hash].push_back(rec);
    ++h->n_traces;
  }
  h->stats.n_gates_lowered += plan.gates.size();
  h->stats.algorithmic_bytes += plan.algorithmic_bytes;
  h->stats.unfused_bytes += plan.unfused_bytes;
  h->perm = plan.perm_out;
  if (h->cfg.world == 1) h->support = plan.support_out;
  return QCB_OK;
}

// bring the state back to the canonical layout (logical bit b at physical bit b)
int restore_layout(qcb_sim* h) {
  if (perm_is_identity(h)) return QCB_OK;
  const int n = h->cfg.n_total, nl = h->cfg.n_local;
  std::vector<int> logical_of(n);
  auto refresh = [&]() { for (int b = 0; b < n; ++b) logical_of[h->perm[b]] = b; };
  refresh();
  // 1) every global physical bit gets its own logical bit back, through a local staging position.  The swaps are worked
  // out on the host first; consecutive ones with pairwise disjoint bits then move in one pass (do_exchange_multi).
  {
    std::vector<std::pair<int, int>> seq;
    for (int G = nl; G < n; ++G) {
      if (logical_of[G] == G) continue;
      int where = h->perm[G];                       // physical position of logical bit G
      if (where >= nl) {                            // sits on another global bit: bring it local first
        int l = nl - 1;
        seq.emplace_back(where, l);
        std::swap(h->perm[logical_of[where]], h->perm[logical_of[l]]);
        refresh();
        where = h->perm[G];
      }
      seq.emplace_back(G, where);
      std::swap(h->perm[logical_of[G]], h->perm[logical_of[where]]);
      refresh();
    }
    std::vector<std::pair<int, int>> batch;
    for (auto& pr : seq) {
      bool clash = false;
      for (auto& b : batch) clash = clash || b.first == pr.first || b.second == pr.second || b.first == pr.second || b.second == pr.first;
      if (clash) { RET(do_exchange_multi(h, batch)); batch.clear(); }
      batch.push_back(pr);
    }
    RET(do_exchange_multi(h, batch));
  }
  // 2) local permutation: swap gates on physical bits (fused into tile sweeps by the scheduler)
  std::vector<Gate> swaps;
  std::vector<int> perm = h->perm;
  std::vector<int> lof(n);
  for (int b = 0; b < n; ++b) lof[perm[b]] = b;
  for (int b = 0; b < nl; ++b) {
    if (perm[b] == b) continue;
    // logical b lives at physical perm[b]; physical b holds logical lof[b]: swap physical bits b and perm[b]
    int pa = b, pb = perm[b];
    Gate g; g.kind = G_SWAPP; g.t0 = std::min(pa, pb); g.t1 = std::max(pa, pb); g.m[0] = {1, 0}; g.frac = 0.5;
    swaps.push_back(g);
    int la = lof[pa], lb = lof[pb];
    std::swap(perm[la], perm[lb]);
    lof[pa] = lb; lof[pb] = la;
  }
  if (!swaps.empty()) {
    // the swap gates are expressed on PHYSICAL bits: schedule them with an identity permutation
    Plan plan; plan.cfg = h->cfg; plan.gates = std::move(swaps);
    RET(ensure_prog(h, std::max<size_t>(1 << 17, plan.gates.size() * 512)));
    if (h->prog_ev_valid) CU(h, cudaEventSynchronize(h->prog_ev));
    StreamingExecutor sink(h);
    int rc = schedule(plan, std::vector<int>(), &sink);
    if (rc == QCB_OK) rc = sink.flush_exchanges();
    CU(h, cudaEventRecord(h->prog_ev, h->stream));
    h->prog_ev_valid = true;
    if (rc != QCB_OK) return plan.error != "stage sink failed" ? fail(h, rc, plan.error) : rc;
  }
  for (int b = 0; b < n; ++b) h->perm[b] = b;
  return QCB_OK;
}

int begin_timing(qcb_sim* h) {
  xtime_resolve(h);                                 // exchanges of earlier calls whose statistics were never read
  std::memset(&h->stats, 0, sizeof h->stats);
  CU(h, cudaEventRecord(h->ev0, h->stream));
  return QCB_OK;
}
int end_timing(qcb_sim* h) {
  CU(h, cudaEventRecord(h->ev1, h->stream));
  h->timing_pending = true;
  return QCB_OK;
}

// normalize-state (domain/state.clj:544-551): divide by ||psi|| when the norm exceeds 1e-12
int normalize_inplace(qcb_sim* h) {
  const int grid = red_grid(h);
  RET(ensure_partials(h, (size_t)grid * 2));
  CU(h, launch_reduce(h->state, h->local_count, 1, h->d_partials, grid, h->stream));
  CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
  RET(allreduce_sum(h, h->d_vals + 8, 2));
  CU(h, launch_finalize(h->d_vals + 8, 1, 2, 2, 1e-12, h->d_vals + 4, h->stream));
  CU(h, launch_scale_dev(h->state, h->local_count, h->d_vals + 4, grid, h->stream));
  h->stats.n_kernel_launches += 4;
  return QCB_OK;
}

int norm_squared(qcb_sim* h, double* out) {
  const int grid = red_grid(h);
  RET(ensure_partials(h, (size_t)grid * 2));
  CU(h, launch_reduce(h->state, h->local_count, 1, h->d_partials, grid, h->stream));
  CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
  RET(allreduce_sum(h, h->d_vals + 8, 2));
  double v[2];
  RET(read_back(h, h->d_vals + 8, sizeof v, v));
  *out = v[0];
  return QCB_OK;
}

// measure-specific-qubits (domain/state.clj:946-1014).  Up to MAX_HIST_BITS measured qubits: one histogram pass.  Up to
// MAX_MEASURE_BITS: two passes - the outcome enumeration (bit i of the outcome index <-> i-th listed qubit, ascending index)
// is major in the bits above MAX_HIST_BITS, so the draw first selects those from their marginal and then the low bits from
// the marginal restricted to that choice: the same "first outcome with cumulative >= r" as one pass over all 2^m outcomes.
int measure_qubits_impl(qcb_sim* h, const int32_t* qubits, int m, double u, int32_t* out_bits, double* out_prob,
                        double* out_probs, bool collapse) {
  const int n = h->cfg.n_total;
  if (m <= 0 || m > MAX_MEASURE_BITS) return fail(h, QCB_ERR_UNSUPPORTED, "measure: 1.." + std::to_string(MAX_MEASURE_BITS) + " qubits per :measure op supported");
  if (out_probs && m > MAX_HIST_BITS) return fail(h, QCB_ERR_UNSUPPORTED, "marginal distribution over more than " + std::to_string(MAX_HIST_BITS) + " qubits: not supported");
  BitList all; all.n = m;
  for (int k = 0; k < m; ++k) {
    if (qubits[k] < 0 || qubits[k] >= n) return fail(h, QCB_ERR_INVALID, "measure: qubit out of range");
    all.pos[k] = h->perm[n - 1 - qubits[k]];
  }
  const uint64_t ext_or = (uint64_t)h->cfg.rank << h->cfg.n_local;
  const int grid = red_grid(h);
  // one histogram pass over `bits` (restricted to filter == fval): probabilities of the 2^bits.n outcomes, summed over ranks
  auto pass = [&](const BitList& bits, const BitList& filter, uint32_t fval, std::vector<double>& probs) -> int {
    const uint32_t nk = 1u << bits.n;
    RET(ensure_partials(h, (size_t)grid * nk + nk));
    CU(h, launch_marginal(h->state, h->local_count, ext_or, bits, filter, fval, h->d_partials, grid, h->stream));
    double* d_out = h->d_partials + (size_t)grid * nk;
    CU(h, launch_finalize(h->d_partials, grid, nk, 0, 0.0, d_out, h->stream));
    RET(allreduce_sum(h, d_out, nk));
    h->stats.n_kernel_launches += 2;
    probs.resize(nk);
    return read_back(h, d_out, nk * sizeof(double), probs.data());
  };
  // first outcome (enumeration order) with cumulative >= r, clamped (state.clj:979-985); returns the cumulative before it
  auto pick = [](const std::vector<double>& probs, double r, double& before) -> uint32_t {
    double cum = 0; uint32_t sel = 0; before = 0;
    for (; sel < probs.size(); ++sel) { if (cum + probs[sel] >= r) break; cum += probs[sel]; }
    if (sel >= probs.size()) { sel = (uint32_t)probs.size() - 1; cum -= probs[sel]; }
    before = cum;
    return sel;
  };
  BitList none; none.n = 0;
  uint32_t sel = 0; double p = 0;
  if (m <= MAX_HIST_BITS) {
    std::vector<double> probs;
    RET(pass(all, none, 0, probs));
    if (out_probs) std::memcpy(out_probs, probs.data(), probs.size() * sizeof(double));
    if (!collapse) return QCB_OK;
    double total = 0;
    for (double v : probs) total += v;
    // r = total * u ; first outcome with cum >= r where cum is the running sum INCLUDING the outcome (the reference's loop)
    const double r = total * u;
    double cum = 0;
    while (sel < probs.size() && (cum += probs[sel]) < r) ++sel;
    if (sel >= probs.size()) sel = (uint32_t)probs.size() - 1;
    p = probs[sel];
  } else {
    BitList lo, hi; lo.n = MAX_HIST_BITS; hi.n = m - MAX_HIST_BITS;
    for (int k = 0; k < lo.n; ++k) lo.pos[k] = all.pos[k];
    for (int k = 0; k < hi.n; ++k) hi.pos[k] = all.pos[MAX_HIST_BITS + k];
    std::vector<double> ph, pl;
    RET(pass(hi, none, 0, ph));
    double total = 0;
    for (double v : ph) total += v;
    const double r = total * u;
    double before = 0;
    const uint32_t kh = pick(ph, r, before);
    RET(pass(lo, hi, kh, pl));
    double b2 = 0;
    const uint32_t kl = pick(pl, r - before, b2);
    sel = kl | (kh << MAX_HIST_BITS);
    p = pl[kl];
  }
  const double factor = p > 0 ? 1.0 / std::sqrt(p) : 1.0;
  CU(h, launch_collapse(h->state, h->local_count, ext_or, all, sel, factor, red_grid(h), h->stream));
  h->stats.n_kernel_launches += 1;
  if (out_bits) for (int k = 0; k < m; ++k) out_bits[k] = (sel >> k) & 1;
  if (out_prob) *out_prob = p;
  return QCB_OK;
}

// apply a list of public ops (splitting at MEASURE ops); draws: optional external stream for MEASURE
int apply_ops_impl(qcb_sim* h, const qcb_op* ops, uint64_t n_ops) {
  uint64_t start = 0;
  auto flush = [&](uint64_t end) -> int {
    if (end <= start) return QCB_OK;
    std::vector<Gate> gates;
    std::string err;
    int rc = lower_ops(h->cfg, ops + start, end - start, gates, err);
    if (rc != QCB_OK) return fail(h, rc, err);
    return run_gates(h, std::move(gates));
  };
  // validate everything first so that a bad op leaves the state untouched (reference: the whole job fails)
  {
    std::vector<Gate> tmp; std::string err; uint64_t s0 = 0;
    for (uint64_t k = 0; k <= n_ops; ++k) {
      if (k == n_ops || ops[k].kind == QCB_OP_MEASURE) {
        if (k > s0) { int rc = lower_ops(h->cfg, ops + s0, k - s0, tmp, err); if (rc != QCB_OK) return fail(h, rc, err); }
        s0 = k + 1;
      }
    }
  }
  for (uint64_t k = 0; k < n_ops; ++k) {
    if (ops[k].kind == QCB_OP_MEASURE) {
      RET(flush(k));
      start = k + 1;
      if (!ops[k].ext) return fail(h, QCB_ERR_INVALID, "Measure requires measurement-qubits parameter");
      RET(measure_qubits_impl(h, static_cast<const int32_t*>(ops[k].ext), ops[k].n_mask, ops[k].angle, nullptr, nullptr, nullptr, true));
    }
  }
  return flush(n_ops);
}

// ---- sampling (domain/state.clj:894-913)
int sample_impl(qcb_sim* h, const double* uniforms, uint64_t n_shots, uint64_t* outcomes) {
  RET(restore_layout(h));
  const uint64_t n_chunks = (h->local_count + SAMPLE_CHUNK - 1) / SAMPLE_CHUNK;
  const size_t off_u = ((n_chunks * 8 + 255) / 256) * 256;
  const size_t off_o = off_u + ((n_shots * 8 + 255) / 256) * 256;
  const size_t off_g = off_o + ((n_shots * 8 + 255) / 256) * 256;
  RET(ensure_scratch(h, off_g + 8 * (size_t)h->cfg.world + 256));
  double* d_chunks = reinterpret_cast<double*>(h->d_scratch);
  double* d_u = reinterpret_cast<double*>(h->d_scratch + off_u);
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(h->d_scratch + off_o);
  double* d_totals = reinterpret_cast<double*>(h->d_scratch + off_g);
  const int grid = (int)std::min<uint64_t>(n_chunks, (uint64_t)h->num_sms * 4);
  CU(h, launch_chunk_sums(h->state, h->local_count, d_chunks, grid, h->stream));
  CU(h, launch_scan_inclusive(d_chunks, n_chunks, h->stream));
  h->stats.n_kernel_launches += 2;
  // totals of all ranks
  std::vector<double> totals(h->cfg.world, 0.0);
  if (h->cfg.world > 1) {
    NC(h, g_nccl.AllGather(d_chunks + (n_chunks - 1), d_totals, 1, ncclDouble, h->comm, h->stream));
    RET(read_back(h, d_totals, 8 * (size_t)h->cfg.world, totals.data()));
  } else {
    RET(read_back(h, d_chunks + (n_chunks - 1), 8, totals.data()));
  }
  double total = 0, offset = 0;
  for (int r = 0; r < h->cfg.world; ++r) { if (r == h->cfg.rank) offset = total; total += totals[r]; }
  if (std::fabs(total - 1.0) > 1e-8)
    return fail(h, QCB_ERR_STATE, "State is not properly normalized: total probability " + std::to_string(total));
  if (n_shots == 0) return QCB_OK;
  RET(ensure_pinned(h, n_shots * 8));
  std::memcpy(h->h_pin, uniforms, n_shots * 8);
  CU(h, cudaMemcpyAsync(d_u, h->h_pin, n_shots * 8, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemsetAsync(d_out, 0, n_shots * 8, h->stream));
  const int sgrid = (int)std::min<uint64_t>(n_shots, 65535);
  CU(h, launch_sample(h->state, h->local_count, d_chunks, n_chunks, d_u, n_shots, total, offset, h->cfg.rank == 0,
                      h->cfg.rank == h->cfg.world - 1, (uint64_t)h->cfg.rank << h->cfg.n_local, d_out, sgrid, h->stream));
  h->stats.n_kernel_launches += 1;
  if (h->cfg.world > 1) NC(h, g_nccl.AllReduce(d_out, d_out, n_shots, ncclUint64, ncclSum, h->comm, h->stream));
  RET(read_back(h, d_out, n_shots * 8, outcomes));
  return QCB_OK;
}

// ---- Pauli / Hamiltonian expectation (domain/observables.clj:216-251, domain/hamiltonian.clj:96-114)
struct PTerm { uint64_t x = 0, z = 0; int ny = 0; size_t idx = 0; };

int expect_terms_impl(qcb_sim* h, const char* const* strings, uint64_t n_terms, double* out_terms) {
  const int n = h->cfg.n_total, nl = h->cfg.n_local;
  const uint64_t local_mask = (nl >= 64) ? ~0ULL : ((1ULL << nl) - 1);
  // Pauli strings as LOGICAL bit masks (char k <-> qubit k <-> bit n-1-k); the physical masks follow the current layout
  struct LTerm { uint64_t x = 0, z = 0; int ny = 0; };
  std::vector<LTerm> lterms(n_terms);
  for (uint64_t t = 0; t < n_terms; ++t) {
    const char* s = strings[t];
    if (!s || (int)std::strlen(s) != n) return fail(h, QCB_ERR_INVALID, "pauli string length must equal the number of qubits");
    for (int q = 0; q < n; ++q) {
      const uint64_t b = 1ULL << (n - 1 - q);
      switch (s[q]) {
        case 'I': break;
        case 'X': lterms[t].x |= b; break;
        case 'Z': lterms[t].z |= b; break;
        case 'Y': lterms[t].x |= b; lterms[t].z |= b; lterms[t].ny++; break;
        default: return fail(h, QCB_ERR_INVALID, "pauli string may contain only I, X, Y, Z");
      }
    }
  }
  auto phys = [&](uint64_t m) { uint64_t r = 0; for (int b = 0; b < n; ++b) if ((m >> b) & 1) r |= 1ULL << h->perm[b]; return r; };
  const int grid = red_grid(h);
  RET(ensure_partials(h, (size_t)grid * EXPECT_TERMS));
  RET(ensure_scratch(h, (n_terms + EXPECT_TERMS) * 8 + 256));
  double* d_res = reinterpret_cast<double*>(h->d_scratch);
  const uint64_t ext_or = (uint64_t)h->cfg.rank << nl;
  static const double PH[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
  // evaluates the given terms (X masks all local): groups of <= 16 terms sharing an X mask, one pass over the state each
  auto evaluate = [&](std::vector<PTerm>& terms) -> int {
    std::stable_sort(terms.begin(), terms.end(), [](const PTerm& a, const PTerm& b) { return a.x < b.x; });
    std::vector<size_t> order;     // result slot -> original term index
    size_t t = 0, slot = 0;
    while (t < terms.size()) {
      ExpectTerms et; et.n = 0;
      const uint64_t x = terms[t].x;
      while (t < terms.size() && terms[t].x == x && et.n < EXPECT_TERMS) {
        et.zmask[et.n] = terms[t].z; et.pr[et.n] = PH[terms[t].ny & 3][0]; et.pi[et.n] = PH[terms[t].ny & 3][1];
        order.push_back(terms[t].idx);
        ++et.n; ++t;
      }
      for (int k = et.n; k < EXPECT_TERMS; ++k) { et.zmask[k] = 0; et.pr[k] = 0; et.pi[k] = 0; }
      const int pivot = x ? 63 - __builtin_clzll(x) : 0;
      CU(h, launch_expect_group(h->state, h->local_count, x, pivot, ext_or, et, h->d_partials, grid, h->stream));
      CU(h, launch_finalize(h->d_partials, grid, EXPECT_TERMS, 0, 0.0, d_res + slot, h->stream));
      h->stats.n_kernel_launches += 2;
      slot += et.n;     // next group overwrites the unused tail of this one
    }
    RET(allreduce_sum(h, d_res, slot));
    std::vector<double> res(slot);
    if (slot) RET(read_back(h, d_res, slot * 8, res.data()));
    for (size_t k = 0; k < slot; ++k) out_terms[order[k]] = res[k];
    return QCB_OK;
  };
  std::vector<char> done(n_terms, 0);
  size_t n_done = 0;
  while (n_done < n_terms) {
    std::vector<PTerm> terms;
    for (uint64_t t = 0; t < n_terms; ++t) {
      if (done[t]) continue;
      const uint64_t px = phys(lterms[t].x);
      if (px & ~local_mask) continue;              // an X / Y factor on a rank bit: its pair partner lives on another GPU
      PTerm pt; pt.x = px; pt.z = phys(lterms[t].z); pt.ny = lterms[t].ny; pt.idx = t;
      terms.push_back(pt);
      done[t] = 1; ++n_done;
    }
    if (!terms.empty()) { RET(evaluate(terms)); continue; }
    // Sharded layout, nothing evaluable: bring the X / Y qubits of the first remaining term into local positions with
    // qubit exchanges (the layout permutation is tracked, nothing needs to be moved back), preferring partners that the
    // remaining terms use least
    uint64_t t0 = 0;
    while (done[t0]) ++t0;
    const uint64_t px0 = phys(lterms[t0].x);
    if (__builtin_popcountll(px0) > nl)
      return fail(h, QCB_ERR_UNSUPPORTED, "pauli string has more X/Y factors than one GPU holds qubits: not supported in the sharded layout");
    for (int g = n - 1; g >= nl; --g) {
      if (!((px0 >> g) & 1)) continue;
      const uint64_t cur = phys(lterms[t0].x);
      int best = -1; size_t best_uses = 0;
      for (int l = nl - 1; l >= 0; --l) {
        if ((cur >> l) & 1) continue;
        size_t uses = 0;
        for (uint64_t t = 0; t < n_terms; ++t) if (!done[t] && ((phys(lterms[t].x) >> l) & 1)) ++uses;
        if (best < 0 || uses < best_uses) { best = l; best_uses = uses; }
        if (l < nl - 8 && best_uses == 0) break;   // a free partner among the high bits is good enough
      }
      if (best < 0) return fail(h, QCB_ERR_UNSUPPORTED, "no local qubit available to localise the pauli string");
      RET(do_exchange(h, g, best));
      std::vector<int> logical_of(n);
      for (int b = 0; b < n; ++b) logical_of[h->perm[b]] = b;
      std::swap(h->perm[logical_of[g]], h->perm[logical_of[best]]);
    }
  }
  return QCB_OK;
}

// Kraus operator K psi / ||K psi|| (domain/channel.clj:162-200).  Operators proportional to a unitary
// are applied pre-scaled (identical up to rounding) to save the norm pass.
bool proportional_to_unitary(const double K[8], double* scale) {
  const double a = K[0] * K[0] + K[1] * K[1] + K[4] * K[4] + K[5] * K[5];   // (K^dagger K)_00
  const double d = K[2] * K[2] + K[3] * K[3] + K[6] * K[6] + K[7] * K[7];   // (K^dagger K)_11
  const double ore = K[0] * K[2] + K[1] * K[3] + K[4] * K[6] + K[5] * K[7]; // (K^dagger K)_01
  const double oim = K[0] * K[3] - K[1] * K[2] + K[4] * K[7] - K[5] * K[6];
  if (a <= 0) return false;
  if (std::fabs(a - d) > 1e-15 * a || std::fabs(ore) > 1e-15 * a || std::fabs(oim) > 1e-15 * a) return false;
  *scale = 1.0 / std::sqrt(a);
  return true;
}

Gate kraus_gate(const qcb_sim* h, const double K[8], int target, double scale) {
  Gate g; g.kind = G_MAT1; g.t0 = h->cfg.n_total - 1 - target; g.frac = 1.0;
  for (int i = 0; i < 4; ++i) g.m[i] = {K[2 * i] * scale, K[2 * i + 1] * scale};
  return g;
}

int noise_target(const qcb_op& op) {
  // noise.clj:70 — (get-in gate [:operation-params :target] 0)
  switch (op.kind) {
    case QCB_OP_I: case QCB_OP_X: case QCB_OP_Y: case QCB_OP_Z: case QCB_OP_H: case QCB_OP_S: case QCB_OP_SDG:
    case QCB_OP_T: case QCB_OP_TDG: case QCB_OP_RX: case QCB_OP_RY: case QCB_OP_RZ: case QCB_OP_PHASE: case QCB_OP_U1Q:
      return op.q[0];
    case QCB_OP_CNOT: case QCB_OP_CZ: case QCB_OP_CY: case QCB_OP_CRX: case QCB_OP_CRY: case QCB_OP_CRZ:
    case QCB_OP_RYDBERG_CZ: case QCB_OP_RYDBERG_CPHASE: case QCB_OP_CU1Q:
      return op.q[1];
    case QCB_OP_TOFFOLI: return op.q[2];
    default: return 0;
  }
}

const qcb_noise_entry* find_noise(const qcb_noise_table* nt, int kind) {
  if (!nt) return nullptr;
  for (int i = 0; i < nt->n_entries; ++i) if (nt->entries[i].op_kind == kind) return &nt->entries[i];
  return nullptr;
}

// channel.clj:225-242 — p_k = max |coeff|^2 ; first k with u < cumulative, else the last
int select_kraus(const qcb_noise_entry* e, double u) {
  double cum = 0;
  for (int k = 0; k < e->n_kraus; ++k) {
    double mx = 0;
    for (int i = 0; i < 4; ++i) mx = std::max(mx, e->kraus[k][2 * i] * e->kraus[k][2 * i] + e->kraus[k][2 * i + 1] * e->kraus[k][2 * i + 1]);
    cum += mx;
    if (u < cum || k >= e->n_kraus - 1) return k;
  }
  return e->n_kraus - 1;
}

void worker_main(qcb_sim* h);
int run_job(qcb_sim* h, Job& job);

}  // namespace

// ---- group handles: fan a call out to the member handles, one host thread per device
template <class F>
static int group_run(qcb_sim* g, F&& f) {
  const int n = (int)g->members.size();
  std::vector<int> rc(n, QCB_OK);
  std::vector<std::thread> th;
  th.reserve(n);
  for (int r = 1; r < n; ++r) th.emplace_back([&, r] { rc[r] = f(g->members[r], r); });
  rc[0] = f(g->members[0], 0);
  for (auto& t : th) t.join();
  for (int r = 0; r < n; ++r)
    if (rc[r] != QCB_OK) { g->err = "device " + std::to_string(g->members[r]->device) + ": " + g->members[r]->err; return rc[r]; }
  return QCB_OK;
}

// =================================================================== C ABI
extern "C" {

int32_t qcb_abi_version(void) { return QCB_ABI_VERSION; }

int32_t qcb_last_error(qcb_handle h, char* buf, size_t len) {
  if (!buf || !len) return QCB_ERR_INVALID;
  std::string m;
  if (h) m = h->err;
  else { std::lock_guard<std::mutex> lk(g_create_mu); m = g_create_error; }
  std::snprintf(buf, len, "%s", m.c_str());
  return QCB_OK;
}

int32_t qcb_device_count(int32_t* count) {
  if (!count) return QCB_ERR_INVALID;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *count = 0; return fail(nullptr, QCB_ERR_CUDA, cudaGetErrorString(e)); }
  *count = c;
  return QCB_OK;
}

int32_t qcb_nccl_unique_id(void* out128) {
  if (!out128) return QCB_ERR_INVALID;
  std::string err;
  if (!load_nccl(err)) return fail(nullptr, QCB_ERR_NCCL, err);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, QCB_ERR_NCCL, g_nccl.GetErrorString(r));
  std::memcpy(out128, &id, sizeof id);
  return QCB_OK;
}

int32_t qcb_config_default(qcb_config* cfg) {
  if (!cfg) return QCB_ERR_INVALID;
  std::memset(cfg, 0, sizeof *cfg);
  cfg->device = -1; cfg->fusion = 1; cfg->strict_parity = 1; cfg->world_size = 1;
  return QCB_OK;
}

// Map the state allocations of ALL other ranks (the multi-qubit swap kernel exchanges with up to 7 partners at once).
// Ranks in other processes: cudaIpcGetMemHandle -> all-gather -> cudaIpcOpenMemHandle.  Ranks of the same process (a group
// handle): direct pointers + cudaDeviceEnablePeerAccess.  Every rank must reach the same verdict (a rank using peer memory
// while its partner sends through NCCL would hang), so the outcome is agreed on with a min-all-reduce; any failure simply
// leaves the NCCL send/recv exchange in place.  QCB_EXCHANGE = swap (default: in-place swap kernel) | ce (copy-engine pull
// through staging buffers, the round-1 path) | nccl (ncclSend/ncclRecv).
static void setup_p2p(qcb_sim* h) {
  const int world = h->cfg.world;
  h->peer_state.assign(world, nullptr);
  const char* mode = getenv("QCB_EXCHANGE");
  const std::string m = mode ? mode : "";
  h->xmode = (m == "nccl") ? 2 : ((m == "ce" || m == "p2p") ? 1 : 0);
  double ok = (h->xmode == 2) ? 0.0 : 1.0;
  if (h->gs) {
    // ---- same process: publish the pointer, wait for everybody, enable peer access
    h->gs->states[h->cfg.rank] = h->state;
    h->gs->devices[h->cfg.rank] = h->device;
    h->gs->agree(true);
    if (ok > 0.0)
      for (int r = 0; r < world; ++r) {
        if (r == h->cfg.rank) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, h->device, h->gs->devices[r]) != cudaSuccess || !can) { cudaGetLastError(); ok = 0.0; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(h->gs->devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0.0; break; }
        cudaGetLastError();
        h->peer_state[r] = h->gs->states[r];
      }
    h->peers_ipc = false;
  } else {
    unsigned char* d_ipc = nullptr;
    std::vector<cudaIpcMemHandle_t> handles(world);
    if (cudaMalloc(&d_ipc, (size_t)world * sizeof(cudaIpcMemHandle_t)) != cudaSuccess) { cudaGetLastError(); return; }
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, h->state) != cudaSuccess) { cudaGetLastError(); ok = 0.0; std::memset(&mine, 0, sizeof mine); }
    cudaMemcpyAsync(d_ipc + (size_t)h->cfg.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream);
    bool comm_ok = g_nccl.AllGather(d_ipc + (size_t)h->cfg.rank * sizeof mine, d_ipc, sizeof mine, ncclChar, h->comm, h->stream) == ncclSuccess;
    comm_ok = comm_ok && cudaMemcpyAsync(handles.data(), d_ipc, (size_t)world * sizeof mine, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
    comm_ok = comm_ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    if (!comm_ok) ok = 0.0;
    if (ok > 0.0)
      for (int peer = 0; peer < world; ++peer) {
        if (peer == h->cfg.rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, handles[peer], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0.0; break; }
        h->peer_state[peer] = static_cast<double2*>(p);
      }
    h->peers_ipc = true;
    cudaFree(d_ipc);
  }
  // agree: peer memory only if every rank mapped all of its partners
  double* d_ok = h->d_vals + 202;
  cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, h->stream);
  if (g_nccl.AllReduce(d_ok, d_ok, 1, ncclDouble, ncclMin, h->comm, h->stream) == ncclSuccess &&
      cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess && cudaStreamSynchronize(h->stream) == cudaSuccess)
    h->p2p = ok > 0.5;
  if (!h->p2p) {
    for (auto& p : h->peer_state) if (p) { if (h->peers_ipc) cudaIpcCloseMemHandle(p); p = nullptr; }
    h->xmode = 2;
  }
}

static int32_t create_group(const qcb_config* c, qcb_handle* out);

// one rank handle on one device; gs != nullptr: member of a single-process group (see create_group)
static int32_t create_single(const qcb_config* c, const std::shared_ptr<GroupShared>& gs, qcb_handle* out) {
  *out = nullptr;
  const int world = c->world_size > 0 ? c->world_size : 1;
  if (world & (world - 1)) return fail(nullptr, QCB_ERR_INVALID, "world_size must be a power of two");
  if (c->rank < 0 || c->rank >= world) return fail(nullptr, QCB_ERR_INVALID, "rank out of range");
  Config cfg = config_from(*c);
  if (cfg.n_total < 1 || cfg.n_total > 62 || cfg.n_local < 1)
    return fail(nullptr, QCB_ERR_INVALID, "n_qubits out of range (need at least 1 local qubit)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, QCB_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  int dev = c->device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) return fail(nullptr, QCB_ERR_INVALID, "device ordinal out of range");
  std::unique_ptr<qcb_sim> h(new qcb_sim());
  h->cfg = cfg; h->device = dev;
  h->gs = gs;
  auto allocate = [&]() -> int {
    qcb_sim* hp = nullptr;   // errors before the handle exists go to the global slot
#define CUC(expr)                                                                                         \
  do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(hp, _e == cudaErrorMemoryAllocation ? QCB_ERR_NOMEM : QCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
    CUC(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, dev));
    h->num_sms = prop.multiProcessorCount;
    CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->local_count = 1ULL << cfg.n_local;
    CUC(cudaMalloc(&h->state, h->local_count * sizeof(double2)));
    CUC(build_tile_maps(h->state, cfg.n_local, &h->maps));
    CUC(cudaMalloc(&h->d_vals, 256 * sizeof(double)));
    CUC(cudaMemsetAsync(h->d_vals, 0, 256 * sizeof(double), h->stream));
    CUC(cudaEventCreate(&h->ev0)); CUC(cudaEventCreate(&h->ev1));
    CUC(cudaEventCreate(&h->xev0)); CUC(cudaEventCreate(&h->xev1));
    CUC(cudaEventCreate(&h->tev0)); CUC(cudaEventCreate(&h->tev1));
    CUC(cudaEventCreateWithFlags(&h->prog_ev, cudaEventDisableTiming));
#undef CUC
    if (world > 1) {
      if (!c->nccl_unique_id) return fail(nullptr, QCB_ERR_INVALID, "world_size > 1 needs nccl_unique_id");
      std::string err;
      if (!load_nccl(err)) return fail(nullptr, QCB_ERR_NCCL, err);
    }
    return QCB_OK;
  };
  int rc = allocate();
  // members of a group agree before anybody enters the (collective) communicator set-up: one failing device must not
  // leave the others waiting inside ncclCommInitRank
  if (gs && !gs->agree(rc == QCB_OK) && rc == QCB_OK) rc = fail(nullptr, QCB_ERR_CUDA, "another device of the group failed to initialise");
  if (rc != QCB_OK) { qcb_destroy(h.release()); return rc; }
  h->perm.resize(cfg.n_total);
  for (int b = 0; b < cfg.n_total; ++b) h->perm[b] = b;
  if (world > 1) {
    ncclUniqueId id;
    std::memcpy(&id, c->nccl_unique_id, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, world, id, c->rank);
    if (r != ncclSuccess) {
      const std::string msg = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r);
      h->comm = nullptr;
      qcb_destroy(h.release());
      return fail(nullptr, QCB_ERR_NCCL, msg);
    }
    setup_p2p(h.get());
  }
  // |0...0>
  cudaMemsetAsync(h->state, 0, h->local_count * sizeof(double2), h->stream);
  if (cfg.rank == 0) launch_set_amp(h->state, 0, 1.0, 0.0, h->stream);
  *out = h.release();
  return QCB_OK;
}

int32_t qcb_create(const qcb_config* c, qcb_handle* out) {
  if (!c || !out) return fail(nullptr, QCB_ERR_INVALID, "null argument");
  *out = nullptr;
  if (c->n_gpus > 1) return create_group(c, out);
  return create_single(c, nullptr, out);
}

static int32_t create_group(const qcb_config* c, qcb_handle* out) {
  const int n = c->n_gpus;
  if (n & (n - 1) || n > 8) return fail(nullptr, QCB_ERR_INVALID, "n_gpus must be 2, 4 or 8");
  if (c->world_size > 1) return fail(nullptr, QCB_ERR_INVALID, "n_gpus > 1 and world_size > 1 are mutually exclusive (one handle for all devices, or one SPMD rank per process)");
  int p = 0;
  while ((1 << p) < n) ++p;
  if (c->n_qubits - p < 1 || c->n_qubits > 62) return fail(nullptr, QCB_ERR_INVALID, "n_qubits out of range (need at least 1 local qubit)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, QCB_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) {
    ids[i] = c->device_ids[i] >= 0 ? c->device_ids[i] : i;
    if (ids[i] >= ndev) return fail(nullptr, QCB_ERR_INVALID, "n_gpus = " + std::to_string(n) + " but device " + std::to_string(ids[i]) + " does not exist (" + std::to_string(ndev) + " visible)");
    for (int j = 0; j < i; ++j) if (ids[j] == ids[i]) return fail(nullptr, QCB_ERR_INVALID, "device_ids must be distinct");
  }
  std::string err;
  if (!load_nccl(err)) return fail(nullptr, QCB_ERR_NCCL, err);
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, QCB_ERR_NCCL, g_nccl.GetErrorString(r));
  std::unique_ptr<qcb_sim> g(new qcb_sim());
  g->is_group = true;
  g->gs = std::make_shared<GroupShared>();
  g->gs->n = n; g->gs->states.assign(n, nullptr); g->gs->devices.assign(n, -1);
  qcb_config single = *c;                       // what the group looks like from outside: one device holding everything
  single.n_gpus = 0; single.world_size = 1; single.rank = 0;
  g->cfg = config_from(single);
  g->local_count = 1ULL << c->n_qubits;
  g->device = ids[0];
  g->members.assign(n, nullptr);
  std::vector<int> rc(n, QCB_OK);
  std::vector<std::string> msgs(n);
  {
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i)
      th.emplace_back([&, i] {
        qcb_config mc = *c;
        mc.n_gpus = 0; mc.world_size = n; mc.rank = i; mc.device = ids[i]; mc.nccl_unique_id = &id;
        rc[i] = create_single(&mc, g->gs, &g->members[i]);
        if (rc[i] != QCB_OK) { std::lock_guard<std::mutex> lk(g_create_mu); msgs[i] = g_create_error; }
      });
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (rc[i] != QCB_OK) {
      const int code = rc[i];
      const std::string msg = "device " + std::to_string(ids[i]) + ": " + msgs[i];
      qcb_destroy(g.release());
      return fail(nullptr, code, msg);
    }
  *out = g.release();
  return QCB_OK;
}

int32_t qcb_destroy(qcb_handle h) {
  if (!h) return QCB_OK;
  {
    std::unique_lock<std::mutex> lk(h->jmu);
    h->stopping = true;
    h->jcv.notify_all();
  }
  if (h->worker_started && h->worker.joinable()) h->worker.join();
  if (h->is_group) {
    // members are torn down concurrently (communicator destruction is collective)
    std::vector<std::thread> th;
    for (qcb_sim* m : h->members) if (m) th.emplace_back([m] { qcb_destroy(m); });
    for (auto& t : th) t.join();
    delete h;
    return QCB_OK;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  xtime_resolve(h);
  for (auto& p : h->peer_state) if (p) { if (h->peers_ipc) cudaIpcCloseMemHandle(p); p = nullptr; }
  if (h->comm) g_nccl.CommDestroy(h->comm);
  tile_prof_dump();
  tile_trace_dump();
  cudaFree(h->state); cudaFree(h->d_prog); cudaFree(h->d_vals); cudaFree(h->d_partials); cudaFree(h->d_scratch); cudaFree(h->xbuf);
  cudaFree(h->noisy_init);
  for (double2* b : h->ckpts) cudaFree(b);
  if (h->h_prog) cudaFreeHost(h->h_prog);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 2; ++i) { if (h->xrecv[i]) cudaEventDestroy(h->xrecv[i]); if (h->xcopy[i]) cudaEventDestroy(h->xcopy[i]); if (h->xpack[i]) cudaEventDestroy(h->xpack[i]); }
  for (cudaEvent_t e : h->xev_pool) cudaEventDestroy(e);
  if (h->xstream) cudaStreamDestroy(h->xstream);
  if (h->xev0) cudaEventDestroy(h->xev0);
  if (h->xev1) cudaEventDestroy(h->xev1);
  if (h->tev0) cudaEventDestroy(h->tev0);
  if (h->tev1) cudaEventDestroy(h->tev1);
  if (h->prog_ev) cudaEventDestroy(h->prog_ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return QCB_OK;
}

// Group handles (qcb_config.n_gpus > 1) dispatch before ENTER: GROUP_ALL runs the same call on every member (identical
// arguments, outputs taken from member 0 where the SPMD path already returns the same result on every rank).
#define GROUP_ALL(h, call_on_m)                                                              \
  if ((h) && (h)->is_group) { std::lock_guard<std::recursive_mutex> _glk((h)->mu); return group_run((h), [&](qcb_sim* m, int r) -> int { (void)r; return (call_on_m); }); }
#define ENTER_KEEP(h)                                          \
  if (!(h)) return QCB_ERR_INVALID;                            \
  std::lock_guard<std::recursive_mutex> _lk((h)->mu);          \
  CU(h, cudaSetDevice((h)->device));
// every entry point but qcb_apply_ops forgets what is known about the support of the state (qcb_sim::support)
#define ENTER(h)                                               \
  ENTER_KEEP(h)                                                \
  (h)->support = ~0ULL;

int32_t qcb_synchronize(qcb_handle h) {
  GROUP_ALL(h, qcb_synchronize(m));
  ENTER(h);
  CU(h, cudaStreamSynchronize(h->stream));
  return QCB_OK;
}

int32_t qcb_set_zero(qcb_handle h) {
  GROUP_ALL(h, qcb_set_zero(m));
  ENTER(h);
  CU(h, cudaMemsetAsync(h->state, 0, h->local_count * sizeof(double2), h->stream));
  if (h->cfg.rank == 0) CU(h, launch_set_amp(h->state, 0, 1.0, 0.0, h->stream));
  for (size_t b = 0; b < h->perm.size(); ++b) h->perm[b] = (int)b;
  // EXPERIMENTAL (off by default, verified on the host emulator only): the sweeps that follow visit only the tiles that can hold
  // non-zero amplitudes (plan.h: Plan::support_in)
  static const bool zero_skip = std::getenv("QCB_ZERO_SKIP") && std::atoi(std::getenv("QCB_ZERO_SKIP")) != 0;
  if (zero_skip && h->cfg.world == 1 && !h->is_group) h->support = 0;
  return QCB_OK;
}

int32_t qcb_set_basis(qcb_handle h, uint64_t index) {
  GROUP_ALL(h, qcb_set_basis(m, index));
  ENTER(h);
  if (h->cfg.n_total < 64 && (index >> h->cfg.n_total)) return fail(h, QCB_ERR_INVALID, "basis index out of range");
  CU(h, cudaMemsetAsync(h->state, 0, h->local_count * sizeof(double2), h->stream));
  if ((index >> h->cfg.n_local) == (uint64_t)h->cfg.rank)
    CU(h, launch_set_amp(h->state, index & (h->local_count - 1), 1.0, 0.0, h->stream));
  for (size_t b = 0; b < h->perm.size(); ++b) h->perm[b] = (int)b;
  return QCB_OK;
}

int32_t qcb_set_state(qcb_handle h, const double* host, uint64_t count) {
  if (h && h->is_group) {      // the whole state: member r takes its slice
    if (!host || count != h->local_count) return fail(h, QCB_ERR_INVALID, "set_state: count must equal 2^n_qubits");
    GROUP_ALL(h, qcb_set_state(m, host + 2 * (uint64_t)r * m->local_count, m->local_count));
  }
  ENTER(h);
  if (!host || count != h->local_count) return fail(h, QCB_ERR_INVALID, "set_state: count must equal the local slice size 2^(n - log2 world)");
  CU(h, cudaMemcpyAsync(h->state, host, count * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  for (size_t b = 0; b < h->perm.size(); ++b) h->perm[b] = (int)b;
  return QCB_OK;
}

// the part of the global range [offset, offset + count) that lies in member r's slice: local offset, count, position in `out`
static void group_slice(const qcb_sim* m, int r, uint64_t offset, uint64_t count, uint64_t& loff, uint64_t& lcnt, uint64_t& opos) {
  const uint64_t lo = (uint64_t)r * m->local_count, hi = lo + m->local_count;
  const uint64_t a = std::max(offset, lo), b = std::min(offset + count, hi);
  if (a >= b) { loff = 0; lcnt = 0; opos = 0; return; }
  loff = a - lo; lcnt = b - a; opos = a - offset;
}

int32_t qcb_get_state(qcb_handle h, uint64_t offset, uint64_t count, double* out) {
  if (h && h->is_group) {      // global range; every member takes part (restoring the canonical layout is collective)
    if (!out || offset + count > h->local_count) return fail(h, QCB_ERR_INVALID, "get_state: range outside the state");
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    return group_run(h, [&](qcb_sim* m, int r) -> int {
      uint64_t lo, lc, op; group_slice(m, r, offset, count, lo, lc, op);
      return qcb_get_state(m, lo, lc, out + 2 * op);
    });
  }
  ENTER(h);
  if (!out || offset + count > h->local_count) return fail(h, QCB_ERR_INVALID, "get_state: range outside the local slice");
  RET(restore_layout(h));
  CU(h, cudaMemcpyAsync(out, h->state + offset, count * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return QCB_OK;
}

int32_t qcb_get_amplitudes(qcb_handle h, const uint64_t* idx, uint64_t n, double* out) {
  if (h && h->is_group) {      // the SPMD path combines across ranks: every member returns the full answer
    if (!idx || !out) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<std::vector<double>> tmp(h->members.size(), std::vector<double>(2 * n));
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_get_amplitudes(m, idx, n, r == 0 ? out : tmp[r].data()); });
    return rc;
  }
  ENTER(h);
  if (!idx || !out) return fail(h, QCB_ERR_INVALID, "null argument");
  if (n == 0) return QCB_OK;
  RET(restore_layout(h));
  std::vector<uint64_t> local(n);
  std::vector<char> mine(n);
  for (uint64_t k = 0; k < n; ++k) {
    if (h->cfg.n_total < 64 && (idx[k] >> h->cfg.n_total)) return fail(h, QCB_ERR_INVALID, "basis index out of range");
    mine[k] = (idx[k] >> h->cfg.n_local) == (uint64_t)h->cfg.rank;
    local[k] = mine[k] ? (idx[k] & (h->local_count - 1)) : 0;
  }
  RET(ensure_scratch(h, n * 8 + n * 16 + 512));
  uint64_t* d_idx = reinterpret_cast<uint64_t*>(h->d_scratch);
  double2* d_out = reinterpret_cast<double2*>(h->d_scratch + ((n * 8 + 255) / 256) * 256);
  CU(h, cudaMemcpyAsync(d_idx, local.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
  CU(h, launch_gather(h->state, d_idx, n, d_out, h->stream));
  std::vector<double> tmp(2 * n);
  RET(read_back(h, d_out, n * 16, tmp.data()));
  for (uint64_t k = 0; k < n; ++k) { out[2 * k] = mine[k] ? tmp[2 * k] : 0.0; out[2 * k + 1] = mine[k] ? tmp[2 * k + 1] : 0.0; }
  if (h->cfg.world > 1) {   // combine across ranks
    CU(h, cudaMemcpyAsync(d_out, out, n * 16, cudaMemcpyHostToDevice, h->stream));
    RET(allreduce_sum(h, reinterpret_cast<double*>(d_out), 2 * n));
    RET(read_back(h, d_out, n * 16, out));
  }
  return QCB_OK;
}

int32_t qcb_normalize(qcb_handle h) {
  GROUP_ALL(h, qcb_normalize(m));
  ENTER(h);
  return normalize_inplace(h);
}

int32_t qcb_state_dev_ptr(qcb_handle h, void** dev_ptr, uint64_t* local_count) {
  if (h && h->is_group) return qcb_state_dev_ptr(h->members[0], dev_ptr, local_count);
  ENTER(h);
  if (dev_ptr) *dev_ptr = h->state;
  if (local_count) *local_count = h->local_count;
  return QCB_OK;
}

int32_t qcb_apply_ops(qcb_handle h, const qcb_op* ops, uint64_t n_ops) {
  GROUP_ALL(h, qcb_apply_ops(m, ops, n_ops));
  ENTER_KEEP(h);                                        // the one entry point that may use a known support (after qcb_set_zero)
  if (!ops && n_ops) return fail(h, QCB_ERR_INVALID, "null ops");
  RET(begin_timing(h));
  h->stats.n_ops = n_ops;
  int rc = apply_ops_impl(h, ops, n_ops);
  end_timing(h);
  return rc;
}

int32_t qcb_norm2(qcb_handle h, double* out) {
  if (h && h->is_group) {
    if (!out) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<double> v(h->members.size(), 0.0);
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_norm2(m, &v[r]); });
    *out = v[0];
    return rc;
  }
  ENTER(h);
  if (!out) return fail(h, QCB_ERR_INVALID, "null argument");
  double s = 0;
  RET(norm_squared(h, &s));
  *out = std::sqrt(s);
  return QCB_OK;
}

int32_t qcb_probabilities(qcb_handle h, uint64_t offset, uint64_t count, double* out) {
  if (h && h->is_group) {
    if (!out || offset + count > h->local_count) return fail(h, QCB_ERR_INVALID, "probabilities: range outside the state");
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    return group_run(h, [&](qcb_sim* m, int r) -> int {
      uint64_t lo, lc, op; group_slice(m, r, offset, count, lo, lc, op);
      if (!lc) { std::lock_guard<std::recursive_mutex> lk(m->mu); if (cudaSetDevice(m->device) != cudaSuccess) return QCB_ERR_CUDA; return restore_layout(m); }
      return qcb_probabilities(m, lo, lc, out + op);
    });
  }
  ENTER(h);
  if (!out || offset + count > h->local_count) return fail(h, QCB_ERR_INVALID, "probabilities: range outside the local slice");
  if (!count) return QCB_OK;
  RET(restore_layout(h));
  // stream through a bounded device buffer so that no 2^n x 8 B scratch is needed next to a 128 GiB state
  const uint64_t chunk = std::min<uint64_t>(count, 1ULL << 24);
  RET(ensure_scratch(h, chunk * 8));
  double* d = reinterpret_cast<double*>(h->d_scratch);
  for (uint64_t done = 0; done < count; done += chunk) {
    const uint64_t c = std::min(chunk, count - done);
    CU(h, launch_probabilities(h->state, offset + done, c, d, (int)std::min<uint64_t>((c + RED_THREADS - 1) / RED_THREADS, (uint64_t)red_grid(h)), h->stream));
    CU(h, cudaMemcpyAsync(out + done, d, c * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
  }
  return QCB_OK;
}

int32_t qcb_sample(qcb_handle h, const double* uniforms, uint64_t n_shots, uint64_t* outcomes) {
  if (h && h->is_group) {
    if (n_shots && (!uniforms || !outcomes)) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<std::vector<uint64_t>> tmp(h->members.size(), std::vector<uint64_t>(n_shots));
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    return group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_sample(m, uniforms, n_shots, r == 0 ? outcomes : tmp[r].data()); });
  }
  ENTER(h);
  if (n_shots && (!uniforms || !outcomes)) return fail(h, QCB_ERR_INVALID, "null argument");
  return sample_impl(h, uniforms, n_shots, outcomes);
}

int32_t qcb_measure_qubits(qcb_handle h, const int32_t* qubits, int32_t m, double u, int32_t* out_bits, double* out_prob) {
  if (h && h->is_group) {
    if (!qubits) return fail(h, QCB_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    return group_run(h, [&](qcb_sim* mb, int r) -> int {
      std::vector<int32_t> bits(m > 0 ? m : 1); double pr = 0;
      return qcb_measure_qubits(mb, qubits, m, u, r == 0 ? out_bits : bits.data(), r == 0 ? out_prob : &pr);
    });
  }
  ENTER(h);
  if (!qubits) return fail(h, QCB_ERR_INVALID, "null argument");
  return measure_qubits_impl(h, qubits, m, u, out_bits, out_prob, nullptr, true);
}

int32_t qcb_marginal_probabilities(qcb_handle h, const int32_t* qubits, int32_t m, double* out_probs) {
  if (h && h->is_group) {
    if (!qubits || !out_probs) return fail(h, QCB_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    return group_run(h, [&](qcb_sim* mb, int r) -> int {
      std::vector<double> pr((size_t)1 << (m > 0 && m < 31 ? m : 0));
      return qcb_marginal_probabilities(mb, qubits, m, r == 0 ? out_probs : pr.data());
    });
  }
  ENTER(h);
  if (!qubits || !out_probs) return fail(h, QCB_ERR_INVALID, "null argument");
  return measure_qubits_impl(h, qubits, m, 0.0, nullptr, nullptr, out_probs, false);
}

int32_t qcb_expect_pauli(qcb_handle h, const char* s, double* out) {
  if (h && h->is_group) {
    if (!s || !out) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<double> v(h->members.size(), 0.0);
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_expect_pauli(m, s, &v[r]); });
    *out = v[0];
    return rc;
  }
  ENTER(h);
  if (!s || !out) return fail(h, QCB_ERR_INVALID, "null argument");
  const char* arr[1] = {s};
  return expect_terms_impl(h, arr, 1, out);
}

int32_t qcb_expect_hamiltonian(qcb_handle h, const double* coeffs, const char* const* strings, uint64_t n_terms,
                               double* out_energy, double* out_terms) {
  if (h && h->is_group) {
    if ((n_terms && (!coeffs || !strings)) || !out_energy) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<double> e(h->members.size(), 0.0);
    std::vector<std::vector<double>> t(h->members.size(), std::vector<double>(n_terms));
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_expect_hamiltonian(m, coeffs, strings, n_terms, &e[r], t[r].data()); });
    *out_energy = e[0];
    if (out_terms && n_terms) std::memcpy(out_terms, t[0].data(), n_terms * 8);
    return rc;
  }
  ENTER(h);
  if ((n_terms && (!coeffs || !strings)) || !out_energy) return fail(h, QCB_ERR_INVALID, "null argument");
  std::vector<double> t(n_terms);
  if (n_terms) RET(expect_terms_impl(h, strings, n_terms, t.data()));
  double e = 0;
  for (uint64_t k = 0; k < n_terms; ++k) e += coeffs[k] * t[k];     // left to right, hamiltonian.clj:108-114
  *out_energy = e;
  if (out_terms) std::memcpy(out_terms, t.data(), n_terms * 8);
  return QCB_OK;
}

int32_t qcb_expect_1q(qcb_handle h, const double mat[8], int32_t target, double* out) {
  if (h && h->is_group) {
    if (!mat || !out) return fail(h, QCB_ERR_INVALID, "bad argument");
    std::vector<double> v(h->members.size(), 0.0);
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_expect_1q(m, mat, target, &v[r]); });
    *out = v[0];
    return rc;
  }
  ENTER(h);
  if (!mat || !out || target < 0 || target >= h->cfg.n_total) return fail(h, QCB_ERR_INVALID, "bad argument");
  int bit = h->perm[h->cfg.n_total - 1 - target];
  if (bit >= h->cfg.n_local) {
    // the pair partner of every amplitude lives on another GPU: bring the qubit into the top local position with one qubit
    // exchange (the layout permutation is tracked; nothing is moved back until a read needs the canonical order)
    const int n = h->cfg.n_total, l = h->cfg.n_local - 1;
    RET(do_exchange(h, bit, l));
    std::vector<int> logical_of(n);
    for (int b = 0; b < n; ++b) logical_of[h->perm[b]] = b;
    std::swap(h->perm[logical_of[bit]], h->perm[logical_of[l]]);
    bit = l;
  }
  Mat2 O; std::memcpy(O.m, mat, sizeof O.m);
  const int grid = red_grid(h);
  RET(ensure_partials(h, (size_t)grid * 2));
  CU(h, launch_expect_1q(h->state, h->local_count, bit, O, h->d_partials, grid, h->stream));
  CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
  RET(allreduce_sum(h, h->d_vals + 8, 2));
  double v[2];
  RET(read_back(h, h->d_vals + 8, sizeof v, v));
  *out = v[0];
  return QCB_OK;
}

int32_t qcb_fidelity(qcb_handle h, const double* host, uint64_t count, double* out) {
  if (h && h->is_group) {
    if (!host || !out || count != h->local_count) return fail(h, QCB_ERR_INVALID, "fidelity: count must equal 2^n_qubits");
    std::vector<double> v(h->members.size(), 0.0);
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_fidelity(m, host + 2 * (uint64_t)r * m->local_count, m->local_count, &v[r]); });
    *out = v[0];
    return rc;
  }
  ENTER(h);
  if (!host || !out || count != h->local_count) return fail(h, QCB_ERR_INVALID, "fidelity: count must equal the local slice size");
  RET(restore_layout(h));
  double2* phi = nullptr;
  cudaError_t e = cudaMalloc(&phi, count * sizeof(double2));
  if (e != cudaSuccess) return fail(h, QCB_ERR_NOMEM, "fidelity: cannot allocate a second state buffer");
  const int grid = red_grid(h);
  int rc = ensure_partials(h, (size_t)grid * 2);
  if (rc == QCB_OK) {
    cudaMemcpyAsync(phi, host, count * sizeof(double2), cudaMemcpyHostToDevice, h->stream);
    // state-fidelity conjugates the FIRST state (the simulator's), domain/state.clj:1176-1185
    launch_inner(h->state, phi, count, h->d_partials, grid, h->stream);
    launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream);
    rc = allreduce_sum(h, h->d_vals + 8, 2);
  }
  double v[2] = {0, 0};
  if (rc == QCB_OK) rc = read_back(h, h->d_vals + 8, sizeof v, v);
  cudaFree(phi);
  if (rc != QCB_OK) return rc;
  *out = std::hypot(v[0], v[1]);
  return QCB_OK;
}

int32_t qcb_apply_kraus_1q(qcb_handle h, const double mat[8], int32_t target) {
  GROUP_ALL(h, qcb_apply_kraus_1q(m, mat, target));
  ENTER(h);
  if (!mat || target < 0 || target >= h->cfg.n_total) return fail(h, QCB_ERR_INVALID, "bad argument");
  RET(begin_timing(h));
  std::vector<Gate> g;
  g.push_back(kraus_gate(h, mat, target, 1.0));
  RET(run_gates(h, std::move(g)));
  int rc = normalize_inplace(h);
  end_timing(h);
  return rc;
}

int32_t qcb_noisy_set_initial_state(qcb_handle h, const double* host, uint64_t count) {
  if (h && h->is_group) return fail(h, QCB_ERR_UNSUPPORTED, "noisy trajectories run as independent replicas per GPU (use one single-GPU handle per device)");
  ENTER(h);
  if (!host) { if (h->noisy_init) { cudaFree(h->noisy_init); h->noisy_init = nullptr; } return QCB_OK; }
  if (h->cfg.world > 1) return fail(h, QCB_ERR_UNSUPPORTED, "noisy trajectories run as independent replicas per GPU (world_size must be 1)");
  if (count != h->local_count) return fail(h, QCB_ERR_INVALID, "noisy initial state: count must equal 2^n_qubits");
  if (!h->noisy_init) {
    cudaError_t e = cudaMalloc(&h->noisy_init, count * sizeof(double2));
    if (e != cudaSuccess) { h->noisy_init = nullptr; return fail(h, QCB_ERR_NOMEM, "noisy initial state: cannot allocate a second state buffer"); }
  }
  CU(h, cudaMemcpyAsync(h->noisy_init, host, count * sizeof(double2), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return QCB_OK;
}

int32_t qcb_noisy_draws_per_shot(qcb_handle h, const qcb_op* ops, uint64_t n_ops, const qcb_noise_table* noise, uint64_t* out) {
  if (!h || !out) return QCB_ERR_INVALID;
  uint64_t cnt = 0;
  for (uint64_t k = 0; k < n_ops; ++k) {
    if (ops[k].kind == QCB_OP_MEASURE) { ++cnt; continue; }
    const qcb_noise_entry* e = find_noise(noise, ops[k].kind);
    if (e && e->n_kraus > 1) ++cnt;
  }
  ++cnt;                                            // final measure-state
  if (noise && noise->has_readout) cnt += (uint64_t)h->cfg.n_total;
  *out = cnt;
  return QCB_OK;
}

int32_t qcb_run_noisy(qcb_handle h, const qcb_op* ops, uint64_t n_ops, const qcb_noise_table* noise, const double* uniforms,
                      uint64_t draws_per_shot, uint64_t n_shots, uint64_t* out_outcomes, double* traj_out, uint64_t max_traj) {
  if (h && h->is_group) return fail(h, QCB_ERR_UNSUPPORTED, "noisy trajectories run as independent replicas per GPU (use one single-GPU handle per device)");
  ENTER(h);
  if ((n_ops && !ops) || (n_shots && (!uniforms || !out_outcomes))) return fail(h, QCB_ERR_INVALID, "null argument");
  if (h->cfg.world > 1) return fail(h, QCB_ERR_UNSUPPORTED, "noisy trajectories run as independent replicas per GPU (world_size must be 1)");
  uint64_t need = 0;
  qcb_noisy_draws_per_shot(h, ops, n_ops, noise, &need);
  if (draws_per_shot < need) return fail(h, QCB_ERR_INVALID, "draws_per_shot too small: need " + std::to_string(need));
  const int n = h->cfg.n_total;
  // validate the circuit once
  {
    std::vector<Gate> tmp; std::string err;
    for (uint64_t k = 0; k < n_ops; ++k) {
      if (ops[k].kind == QCB_OP_MEASURE) continue;
      int rc = lower_ops(h->cfg, ops + k, 1, tmp, err);
      if (rc != QCB_OK) return fail(h, rc, err);
    }
  }
  RET(begin_timing(h));
  h->stats.n_ops = n_ops * n_shots;
  // ---- trajectory tree (default for states up to 1 GiB): shots that share a history share the work.  The circuit is cut
  // into segments at its mid-circuit :measure ops.  Inside a segment the Kraus operator applied after a noisy gate is chosen
  // by a draw and a state-independent probability table (channel.clj:225-233), so the shots of a node are grouped by their
  // choice tuple for the whole segment and every group runs the segment as one fused plan; at a :measure the group's shots
  // split by the outcome their draw selects from the marginal of the shared state (state.clj:946-1014).  States are
  // checkpointed on the device where a node has several children.  Without :measure ops this is exactly "one evolution per
  // distinct Kraus sequence"; with them it replaces the shot-by-shot loop.  Same draws -> same outcomes, trajectories and
  // final state as that loop.
  {
    static const bool tree_off = std::getenv("QCB_NOISY_TREE") && std::atoi(std::getenv("QCB_NOISY_TREE")) == 0;
    bool tree_ok = !tree_off && n <= 26 && n_shots > 0;
    for (uint64_t k = 0; k < n_ops && tree_ok; ++k) if (ops[k].kind == QCB_OP_MEASURE && (ops[k].n_mask > MAX_HIST_BITS || !ops[k].ext)) tree_ok = false;
    if (tree_ok) {
      std::vector<double2*>& ckpts = h->ckpts;          // pool of device checkpoints (depth of the open splits), kept by the handle
      size_t ckpt_used = 0;
      auto ckpt_get = [&](double2** out) -> int {
        if (ckpt_used == ckpts.size()) {
          double2* b = nullptr;
          if (cudaMalloc(&b, h->local_count * sizeof(double2)) != cudaSuccess) { cudaGetLastError(); return fail(h, QCB_ERR_NOMEM, "noisy trajectories: out of device memory for a state checkpoint"); }
          ckpts.push_back(b);
        }
        *out = ckpts[ckpt_used++];
        return QCB_OK;
      };
      const uint64_t last_shot = n_shots - 1;
      std::vector<double> us;
      std::vector<uint64_t> outs;
      auto flush = [&](std::vector<Gate>& pending) -> int {
        if (pending.empty()) return QCB_OK;
        std::vector<Gate> g; g.swap(pending);
        return run_gates(h, std::move(g));
      };
      // the Kraus operator `kidx` of entry e after the gate on qubit tq: pre-scaled when proportional to a unitary, else
      // applied and followed by the norm pass (channel.clj:162-200)
      auto push_kraus = [&](std::vector<Gate>& pending, const qcb_noise_entry* e, int kidx, int tq) -> int {
        double scale = 1.0;
        if (proportional_to_unitary(e->kraus[kidx], &scale)) { pending.push_back(kraus_gate(h, e->kraus[kidx], tq, scale)); return QCB_OK; }
        pending.push_back(kraus_gate(h, e->kraus[kidx], tq, 1.0));
        RET(flush(pending));
        return normalize_inplace(h);
      };
      // splits `shots` by key, the group that holds the run's last shot last (the handle keeps that shot's final state)
      auto split = [&](const std::vector<uint64_t>& shots, const std::vector<uint32_t>& keys, std::vector<std::pair<uint32_t, std::vector<uint64_t>>>& groups) {
        std::map<uint32_t, size_t> index;
        for (size_t j = 0; j < shots.size(); ++j) {
          auto it = index.find(keys[j]);
          if (it == index.end()) { it = index.emplace(keys[j], groups.size()).first; groups.emplace_back(keys[j], std::vector<uint64_t>()); }
          groups[it->second].second.push_back(shots[j]);
        }
        for (size_t g = 0; g + 1 < groups.size(); ++g)
          if (std::find(groups[g].second.begin(), groups[g].second.end(), last_shot) != groups[g].second.end()) { std::swap(groups[g], groups.back()); break; }
      };
      // One SEGMENT = the ops up to the next :measure (or the end).  The Kraus choices inside a segment do not depend on the
      // state, so the shots of a node are grouped by their whole choice tuple for the segment up front and each group runs
      // the segment as ONE fused plan (a split at every noisy gate would cut the plan into one sweep per gate).  The tree
      // branches where it must: per distinct choice tuple at a segment start, per outcome at a :measure.
      std::function<int(uint64_t, uint64_t, std::vector<uint64_t>&)> walk =
          [&](uint64_t k0, uint64_t di0, std::vector<uint64_t>& shots) -> int {
        uint64_t k1 = k0;
        while (k1 < n_ops && ops[k1].kind != QCB_OP_MEASURE) ++k1;
        // multi-Kraus gates of the segment, in order: (op index, entry)
        std::vector<std::pair<uint64_t, const qcb_noise_entry*>> multi;
        for (uint64_t k = k0; k < k1; ++k) {
          const qcb_noise_entry* e = find_noise(noise, ops[k].kind);
          if (e && e->n_kraus > 1) multi.emplace_back(k, e);
        }
        const uint64_t di1 = di0 + multi.size();                 // draw index after the segment
        // group the shots by their choice tuple
        std::vector<std::pair<std::string, std::vector<uint64_t>>> groups;
        {
          std::map<std::string, size_t> index;
          std::string sig(multi.size(), '\0');
          for (uint64_t sh : shots) {
            for (size_t c = 0; c < multi.size(); ++c) sig[c] = (char)select_kraus(multi[c].second, uniforms[sh * draws_per_shot + di0 + c]);
            auto it = index.find(sig);
            if (it == index.end()) { it = index.emplace(sig, groups.size()).first; groups.emplace_back(sig, std::vector<uint64_t>()); }
            groups[it->second].second.push_back(sh);
          }
          for (size_t g = 0; g + 1 < groups.size(); ++g)
            if (std::find(groups[g].second.begin(), groups[g].second.end(), last_shot) != groups[g].second.end()) { std::swap(groups[g], groups.back()); break; }
        }
        double2* ck = nullptr;
        if (groups.size() > 1) { RET(ckpt_get(&ck)); CU(h, cudaMemcpyAsync(ck, h->state, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream)); }
        for (size_t g = 0; g < groups.size(); ++g) {
          if (g) CU(h, cudaMemcpyAsync(h->state, ck, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
          std::vector<uint64_t>& gs = groups[g].second;
          // ---- the segment with this group's choices, as one plan (cut only where a non-unitary Kraus operator needs the norm)
          std::vector<Gate> pending;
          std::string err;
          size_t c = 0;
          for (uint64_t k = k0; k < k1; ++k) {
            const qcb_op& op = ops[k];
            int rc = lower_ops(h->cfg, &op, 1, pending, err);
            if (rc != QCB_OK) return fail(h, rc, err);
            const qcb_noise_entry* e = find_noise(noise, op.kind);
            if (!e || e->n_kraus < 1) continue;
            int tq = noise_target(op);
            if (tq < 0 || tq >= n) tq = 0;
            const int kidx = (e->n_kraus == 1) ? 0 : (int)groups[g].first[c++];
            RET(push_kraus(pending, e, kidx, tq));
          }
          RET(flush(pending));
          if (k1 == n_ops) {
            // ---- leaf: every shot of the group shares the final state.  One measure-state draw each, then readout noise
            // (noise.clj:193-202)
            us.resize(gs.size());
            outs.assign(gs.size(), 0);
            for (size_t j = 0; j < gs.size(); ++j) us[j] = uniforms[gs[j] * draws_per_shot + di1];
            RET(sample_impl(h, us.data(), gs.size(), outs.data()));
            for (size_t j = 0; j < gs.size(); ++j) {
              const double* uj = uniforms + gs[j] * draws_per_shot;
              uint64_t dj = di1 + 1, outcome = outs[j];
              if (noise && noise->has_readout) {
                std::vector<int> flipped;
                for (int q = 0; q < n; ++q) {
                  const int bitpos = n - 1 - q;
                  const int orig = (outcome >> bitpos) & 1;
                  double factor = 1.0;
                  if (noise->correlation) for (int src : flipped) factor *= noise->correlation[(size_t)src * n + q];
                  double eff = (orig ? noise->prob_1_to_0 : noise->prob_0_to_1) * factor;
                  eff = std::min(1.0, std::max(0.0, eff));
                  if (uj[dj++] < eff) { outcome ^= 1ULL << bitpos; flipped.push_back(q); }
                }
              }
              out_outcomes[gs[j]] = outcome;
              if (traj_out && gs[j] < max_traj)
                CU(h, cudaMemcpyAsync(traj_out + gs[j] * 2 * h->local_count, h->state, h->local_count * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
            }
            continue;
          }
          // ---- :measure at k1: the group's shots split by the outcome their draw selects from the shared state's marginal
          const qcb_op& mop = ops[k1];
          const int32_t* mq = static_cast<const int32_t*>(mop.ext);
          const int m = mop.n_mask;
          std::vector<double> probs((size_t)1 << m);
          RET(measure_qubits_impl(h, mq, m, 0.0, nullptr, nullptr, probs.data(), false));
          double total = 0;
          for (double v : probs) total += v;
          std::vector<uint32_t> keys(gs.size());
          for (size_t j = 0; j < gs.size(); ++j) {             // the selection rule of measure_qubits_impl, per shot
            const double r = total * uniforms[gs[j] * draws_per_shot + di1];
            uint32_t sel = 0; double cum = 0;
            while (sel < probs.size() && (cum += probs[sel]) < r) ++sel;
            keys[j] = sel >= probs.size() ? (uint32_t)probs.size() - 1 : sel;
          }
          std::vector<std::pair<uint32_t, std::vector<uint64_t>>> sub;
          split(gs, keys, sub);
          double2* ck2 = nullptr;
          if (sub.size() > 1) { RET(ckpt_get(&ck2)); CU(h, cudaMemcpyAsync(ck2, h->state, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream)); }
          for (size_t q = 0; q < sub.size(); ++q) {
            if (q) CU(h, cudaMemcpyAsync(h->state, ck2, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
            // collapse onto the sub-group's outcome: any of its draws selects it
            RET(measure_qubits_impl(h, mq, m, uniforms[sub[q].second[0] * draws_per_shot + di1], nullptr, nullptr, nullptr, true));
            RET(walk(k1 + 1, di1 + 1, sub[q].second));
          }
          if (sub.size() > 1) --ckpt_used;
        }
        if (groups.size() > 1) --ckpt_used;
        return QCB_OK;
      };
      if (h->noisy_init) {
        CU(h, cudaMemcpyAsync(h->state, h->noisy_init, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
      } else {
        CU(h, cudaMemsetAsync(h->state, 0, h->local_count * sizeof(double2), h->stream));
        CU(h, launch_set_amp(h->state, 0, 1.0, 0.0, h->stream));
      }
      for (size_t b = 0; b < h->perm.size(); ++b) h->perm[b] = (int)b;
      std::vector<uint64_t> all(n_shots);
      for (uint64_t sidx = 0; sidx < n_shots; ++sidx) all[sidx] = sidx;
      int rc = walk(0, 0, all);
      cudaStreamSynchronize(h->stream);
      if (rc != QCB_OK) return rc;
      end_timing(h);
      return QCB_OK;
    }
  }
  // ---- fallback (QCB_NOISY_TREE=0, states above 1 GiB, :measure of more than 12 qubits)
  // The Kraus operator applied after a noisy gate is chosen by a draw and a state-independent probability table
  // (channel.clj:225-233), so without mid-circuit :measure ops the final state of a shot is a function of its sequence of
  // choices only.  Shots are grouped by that sequence: one state evolution per distinct sequence, then all of the group's
  // final measure-state draws go through one sampler call.  Same draws -> same outcomes as the shot-by-shot loop.
  bool has_measure = false;
  std::vector<uint64_t> choice_draw;                 // draw index of every multi-Kraus choice
  std::vector<const qcb_noise_entry*> choice_entry;
  {
    uint64_t di = 0;
    for (uint64_t k = 0; k < n_ops; ++k) {
      if (ops[k].kind == QCB_OP_MEASURE) { has_measure = true; ++di; }
      const qcb_noise_entry* e = find_noise(noise, ops[k].kind);
      if (e && e->n_kraus > 1) { choice_draw.push_back(di++); choice_entry.push_back(e); }
    }
  }
  static const bool no_grouping = std::getenv("QCB_NOISY_GROUPING") && std::atoi(std::getenv("QCB_NOISY_GROUPING")) == 0;
  std::vector<std::vector<uint64_t>> groups;         // shot indices per distinct choice sequence
  if (has_measure || no_grouping) {
    groups.resize(n_shots);
    for (uint64_t s = 0; s < n_shots; ++s) groups[s].push_back(s);
  } else {
    std::map<std::string, size_t> index;
    std::string sig(choice_draw.size(), '\0');
    for (uint64_t s = 0; s < n_shots; ++s) {
      const double* u = uniforms + s * draws_per_shot;
      for (size_t c = 0; c < choice_draw.size(); ++c) sig[c] = (char)select_kraus(choice_entry[c], u[choice_draw[c]]);
      auto it = index.find(sig);
      if (it == index.end()) { it = index.emplace(sig, groups.size()).first; groups.emplace_back(); }
      groups[it->second].push_back(s);
    }
    // the handle keeps the LAST shot's state (:final-state, hardware_simulator.clj:150-160): its group runs last
    if (n_shots) std::swap(groups[index[sig]], groups.back());
  }
  std::vector<double> us;
  std::vector<uint64_t> outs;
  for (const auto& grp : groups) {
    const double* u = uniforms + grp[0] * draws_per_shot;
    uint64_t di = 0;
    if (h->noisy_init) {      // (or (:initial-state options) zero-state), hardware_simulator.clj:128-131
      CU(h, cudaMemcpyAsync(h->state, h->noisy_init, h->local_count * sizeof(double2), cudaMemcpyDeviceToDevice, h->stream));
    } else {
      CU(h, cudaMemsetAsync(h->state, 0, h->local_count * sizeof(double2), h->stream));
      CU(h, launch_set_amp(h->state, 0, 1.0, 0.0, h->stream));
    }
    for (size_t b = 0; b < h->perm.size(); ++b) h->perm[b] = (int)b;
    std::vector<Gate> pending;
    std::string err;
    auto flush = [&]() -> int { if (pending.empty()) return QCB_OK; std::vector<Gate> g; g.swap(pending); return run_gates(h, std::move(g)); };
    for (uint64_t k = 0; k < n_ops; ++k) {
      const qcb_op& op = ops[k];
      if (op.kind == QCB_OP_MEASURE) {
        RET(flush());
        RET(measure_qubits_impl(h, static_cast<const int32_t*>(op.ext), op.n_mask, u[di++], nullptr, nullptr, nullptr, true));
      } else {
        int rc = lower_ops(h->cfg, &op, 1, pending, err);
        if (rc != QCB_OK) return fail(h, rc, err);
      }
      const qcb_noise_entry* e = find_noise(noise, op.kind);
      if (!e || e->n_kraus < 1) continue;
      int tq = noise_target(op);
      if (tq < 0 || tq >= n) tq = 0;
      const int kidx = (e->n_kraus == 1) ? 0 : select_kraus(e, u[di++]);
      double scale = 1.0;
      if (proportional_to_unitary(e->kraus[kidx], &scale)) {
        pending.push_back(kraus_gate(h, e->kraus[kidx], tq, scale));
      } else {
        pending.push_back(kraus_gate(h, e->kraus[kidx], tq, 1.0));
        RET(flush());
        RET(normalize_inplace(h));
      }
    }
    RET(flush());
    // final measurement of every shot of the group (one measure-state draw each), then readout noise (noise.clj:193-202)
    us.resize(grp.size());
    outs.assign(grp.size(), 0);
    for (size_t j = 0; j < grp.size(); ++j) us[j] = uniforms[grp[j] * draws_per_shot + di];
    RET(sample_impl(h, us.data(), grp.size(), outs.data()));
    for (size_t j = 0; j < grp.size(); ++j) {
      const double* uj = uniforms + grp[j] * draws_per_shot;
      uint64_t dj = di + 1;
      uint64_t outcome = outs[j];
      if (noise && noise->has_readout) {
        std::vector<int> flipped;
        for (int q = 0; q < n; ++q) {
          const int bitpos = n - 1 - q;
          const int orig = (outcome >> bitpos) & 1;
          double factor = 1.0;
          if (noise->correlation) for (int src : flipped) factor *= noise->correlation[(size_t)src * n + q];
          double eff = (orig ? noise->prob_1_to_0 : noise->prob_0_to_1) * factor;
          eff = std::min(1.0, std::max(0.0, eff));
          if (uj[dj++] < eff) { outcome ^= 1ULL << bitpos; flipped.push_back(q); }
        }
      }
      out_outcomes[grp[j]] = outcome;
      if (traj_out && grp[j] < max_traj)
        CU(h, cudaMemcpyAsync(traj_out + grp[j] * 2 * h->local_count, h->state, h->local_count * sizeof(double2), cudaMemcpyDeviceToHost, h->stream));
    }
    // the copies above read h->state, which the next group overwrites: stream order keeps them correct
  }
  CU(h, cudaStreamSynchronize(h->stream));
  end_timing(h);
  return QCB_OK;
}

int32_t qcb_get_stats(qcb_handle h, qcb_stats* out) {
  if (h && h->is_group) {      // member 0's counters; device times are the maximum over the members
    if (!out) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<qcb_stats> st(h->members.size());
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_get_stats(m, &st[r]); });
    if (rc != QCB_OK) return rc;
    *out = st[0];
    for (auto& q : st) { out->gpu_ms = std::max(out->gpu_ms, q.gpu_ms); out->exchange_ms = std::max(out->exchange_ms, q.exchange_ms); }
    return QCB_OK;
  }
  ENTER(h);
  if (!out) return fail(h, QCB_ERR_INVALID, "null argument");
  if (h->timing_pending) {
    CU(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    h->stats.gpu_ms = ms;
    h->timing_pending = false;
  }
  xtime_resolve(h);
  *out = h->stats;
  return QCB_OK;
}

int32_t qcb_timer_start(qcb_handle h) {
  GROUP_ALL(h, qcb_timer_start(m));
  ENTER(h);
  CU(h, cudaEventRecord(h->tev0, h->stream));
  return QCB_OK;
}

int32_t qcb_timer_stop(qcb_handle h, double* out_ms) {
  if (h && h->is_group) {
    if (!out_ms) return fail(h, QCB_ERR_INVALID, "null argument");
    std::vector<double> v(h->members.size(), 0.0);
    std::lock_guard<std::recursive_mutex> _glk(h->mu);
    int rc = group_run(h, [&](qcb_sim* m, int r) -> int { return qcb_timer_stop(m, &v[r]); });
    *out_ms = *std::max_element(v.begin(), v.end());
    return rc;
  }
  ENTER(h);
  if (!out_ms) return fail(h, QCB_ERR_INVALID, "null argument");
  CU(h, cudaEventRecord(h->tev1, h->stream));
  CU(h, cudaEventSynchronize(h->tev1));
  float ms = 0;
  CU(h, cudaEventElapsedTime(&ms, h->tev0, h->tev1));
  *out_ms = ms;
  return QCB_OK;
}

// ---- host-only planning API
struct qcb_plan { Plan plan; };

int32_t qcb_plan_create(const qcb_config* cfg, const qcb_op* ops, uint64_t n_ops, qcb_plan** out) {
  if (!cfg || !out || (!ops && n_ops)) return fail(nullptr, QCB_ERR_INVALID, "null argument");
  std::unique_ptr<qcb_plan> p(new qcb_plan());
  p->plan.cfg = config_from(*cfg);
  if (p->plan.cfg.n_total < 1 || p->plan.cfg.n_local < 1) return fail(nullptr, QCB_ERR_INVALID, "n_qubits out of range");
  std::string err;
  int rc = lower_ops(p->plan.cfg, ops, n_ops, p->plan.gates, err);
  if (rc != QCB_OK) return fail(nullptr, rc, err);
  rc = schedule(p->plan, std::vector<int>());
  if (rc != QCB_OK) return fail(nullptr, rc, p->plan.error);
  *out = p.release();
  return QCB_OK;
}
int32_t qcb_plan_create_replayed(const qcb_config* cfg, const qcb_op* ops_recorded, const qcb_op* ops, uint64_t n_ops, qcb_plan** out) {
  if (!cfg || !out || ((!ops || !ops_recorded) && n_ops)) return fail(nullptr, QCB_ERR_INVALID, "null argument");
  const Config c = config_from(*cfg);
  if (c.n_total < 1 || c.n_local < 1) return fail(nullptr, QCB_ERR_INVALID, "n_qubits out of range");
  std::string err;
  Plan first; first.cfg = c;
  int rc = lower_ops(c, ops_recorded, n_ops, first.gates, err);
  if (rc != QCB_OK) return fail(nullptr, rc, err);
  PlanTrace trace;
  rc = schedule(first, std::vector<int>(), nullptr, &trace, nullptr);
  if (rc != QCB_OK) return fail(nullptr, rc, first.error);
  std::unique_ptr<qcb_plan> p(new qcb_plan());
  p->plan.cfg = c;
  rc = lower_ops(c, ops, n_ops, p->plan.gates, err);
  if (rc != QCB_OK) return fail(nullptr, rc, err);
  std::vector<uint64_t> key;
  plan_structure_key(c, p->plan.gates, std::vector<int>(), key);
  if (key != trace.key) return fail(nullptr, QCB_ERR_INVALID, "the two op lists differ in structure");
  rc = schedule(p->plan, std::vector<int>(), nullptr, nullptr, &trace);
  if (rc != QCB_OK) return fail(nullptr, rc, p->plan.error);
  *out = p.release();
  return QCB_OK;
}
int32_t qcb_plan_destroy(qcb_plan* p) { delete p; return QCB_OK; }
int32_t qcb_plan_serialize(const qcb_plan* p, uint64_t* out_words, uint64_t capacity, uint64_t* n_words) {
  if (!p || !n_words) return QCB_ERR_INVALID;
  *n_words = p->plan.words.size();
  if (out_words) std::memcpy(out_words, p->plan.words.data(), 8 * std::min<uint64_t>(capacity, p->plan.words.size()));
  return QCB_OK;
}
int32_t qcb_plan_summary(const qcb_plan* p, uint64_t* n_stages, uint64_t* n_rounds, uint64_t* n_exchanges) {
  if (!p) return QCB_ERR_INVALID;
  if (n_stages) *n_stages = p->plan.stages.size();
  if (n_rounds) *n_rounds = p->plan.n_rounds;
  if (n_exchanges) *n_exchanges = p->plan.n_exchanges;
  return QCB_OK;
}

// ---- jobs (adapter/backend/ideal_simulator.clj:100-176)
int32_t qcb_submit(qcb_handle h, const qcb_job_request* req, uint64_t* out_job_id) {
  if (!h || !req || !out_job_id) return QCB_ERR_INVALID;
  auto job = std::make_shared<Job>();
  job->ops.assign(req->ops, req->ops + req->n_ops);
  for (auto& op : job->ops) {           // own the ext payloads: the caller's pointers die after this call
    if (!op.ext) continue;
    if (op.kind == QCB_OP_U2Q) { job->ext_d.emplace_back(static_cast<const double*>(op.ext), static_cast<const double*>(op.ext) + 32); op.ext = nullptr; }
    else if (op.kind == QCB_OP_MEASURE) { job->ext_i.emplace_back(static_cast<const int32_t*>(op.ext), static_cast<const int32_t*>(op.ext) + op.n_mask); op.ext = nullptr; }
  }
  { size_t di = 0, ii = 0;
    for (auto& op : job->ops) {
      if (op.kind == QCB_OP_U2Q && di < job->ext_d.size()) op.ext = job->ext_d[di++].data();
      else if (op.kind == QCB_OP_MEASURE && ii < job->ext_i.size()) op.ext = job->ext_i[ii++].data();
    } }
  if (req->initial_state && req->initial_count) job->initial.assign(req->initial_state, req->initial_state + 2 * req->initial_count);
  if (req->uniforms && req->n_shots) job->uniforms.assign(req->uniforms, req->uniforms + req->n_shots);
  for (uint64_t t = 0; t < req->n_terms; ++t) { job->ham_coeffs.push_back(req->ham_coeffs[t]); job->ham_strings.emplace_back(req->ham_strings[t]); }
  job->want_probs = req->want_probabilities; job->want_state = req->want_state;
  bool inline_run;
  {
    std::unique_lock<std::mutex> lk(h->jmu);
    job->id = h->next_job++;
    h->jobs[job->id] = job;
    // small jobs finish inside submit so that the caller's first status poll already sees :completed
    // (the reference's blocking helper sleeps 100 ms between polls, application/backend.clj:255)
    inline_run = h->cfg.n_total <= 20 && h->queue.empty();
    if (!inline_run) {
      h->queue.push_back(job);
      if (!h->worker_started) { h->worker = std::thread(worker_main, h); h->worker_started = true; }
      h->jcv.notify_all();
    }
  }
  *out_job_id = job->id;
  if (inline_run) {
    std::lock_guard<std::recursive_mutex> lk(h->mu);
    run_job(h, *job);
  }
  return QCB_OK;
}

int32_t qcb_job_status(qcb_handle h, uint64_t id, int32_t* out_status) {
  if (!h || !out_status) return QCB_ERR_INVALID;
  std::unique_lock<std::mutex> lk(h->jmu);
  auto it = h->jobs.find(id);
  if (it == h->jobs.end()) { *out_status = QCB_JOB_NOT_FOUND; return QCB_OK; }
  *out_status = it->second->status.load();
  return QCB_OK;
}

int32_t qcb_job_result_get(qcb_handle h, uint64_t id, qcb_job_result* r) {
  if (!h || !r) return QCB_ERR_INVALID;
  std::shared_ptr<Job> job;
  {
    std::unique_lock<std::mutex> lk(h->jmu);
    auto it = h->jobs.find(id);
    if (it == h->jobs.end()) { r->status = QCB_JOB_NOT_FOUND; return QCB_ERR_NOTFOUND; }
    job = it->second;
  }
  r->status = job->status.load();
  r->execution_time_ms = job->exec_ms;
  std::snprintf(r->error_message, sizeof r->error_message, "%s", job->error.c_str());
  if (r->status != QCB_JOB_COMPLETED) return QCB_OK;
  if (job->payload_dropped) {
    std::snprintf(r->error_message, sizeof r->error_message, "result payload released (only the 64 most recently finished jobs keep theirs)");
    r->n_shots = 0; r->energy = job->energy; r->has_energy = job->has_energy;
    return QCB_OK;
  }
  if (r->outcomes) std::memcpy(r->outcomes, job->outcomes.data(), 8 * std::min<uint64_t>(r->n_shots, job->outcomes.size()));
  r->n_shots = job->outcomes.size();
  r->energy = job->energy; r->has_energy = job->has_energy;
  if (r->probabilities) std::memcpy(r->probabilities, job->probs.data(), 8 * std::min<uint64_t>(r->prob_capacity, job->probs.size()));
  if (r->state) std::memcpy(r->state, job->state.data(), 8 * std::min<uint64_t>(2 * r->state_capacity, job->state.size()));
  return QCB_OK;
}

int32_t qcb_cancel(qcb_handle h, uint64_t id, int32_t* out_status) {
  if (!h) return QCB_ERR_INVALID;
  std::unique_lock<std::mutex> lk(h->jmu);
  auto it = h->jobs.find(id);
  if (it == h->jobs.end()) { if (out_status) *out_status = QCB_JOB_NOT_FOUND; return QCB_OK; }
  int st = it->second->status.load();
  if (st == QCB_JOB_QUEUED || st == QCB_JOB_RUNNING) {
    it->second->cancel.store(true);          // a running job stops between kernel stages
    if (st == QCB_JOB_QUEUED) it->second->status.store(QCB_JOB_CANCELLED);
    if (out_status) *out_status = QCB_JOB_CANCELLED;
  } else if (out_status) {
    *out_status = st;                        // :cannot-cancel: already finished
  }
  return QCB_OK;
}

int32_t qcb_job_release(qcb_handle h, uint64_t id) {
  if (!h) return QCB_ERR_INVALID;
  std::unique_lock<std::mutex> lk(h->jmu);
  auto it = h->jobs.find(id);
  if (it == h->jobs.end()) return QCB_ERR_NOTFOUND;
  const int st = it->second->status.load();
  if (st == QCB_JOB_QUEUED || st == QCB_JOB_RUNNING) return fail(h, QCB_ERR_STATE, "job is still queued or running: cancel it first");
  h->jobs.erase(it);
  for (auto f = h->finished.begin(); f != h->finished.end(); ++f) if (*f == id) { h->finished.erase(f); break; }
  return QCB_OK;
}

int32_t qcb_queue_status(qcb_handle h, uint64_t* queued, uint64_t* running, uint64_t* completed) {
  if (!h) return QCB_ERR_INVALID;
  std::unique_lock<std::mutex> lk(h->jmu);
  uint64_t q = 0, r = 0, c = 0;
  for (auto& kv : h->jobs) {
    int st = kv.second->status.load();
    q += st == QCB_JOB_QUEUED; r += st == QCB_JOB_RUNNING; c += st == QCB_JOB_COMPLETED;
  }
  if (queued) *queued = q;
  if (running) *running = r;
  if (completed) *completed = c;
  return QCB_OK;
}

// ---- P2: small dense linear algebra on the GPU (domain/math/protocols.clj MatrixAlgebra subset)
static int la_run(qcb_handle h, const std::vector<std::pair<const double*, uint64_t>>& ins, uint64_t out_count, double* out,
                  const std::function<cudaError_t(std::vector<double2*>&, double2*)>& body) {
  size_t total = 0;
  std::vector<size_t> offs;
  for (auto& in : ins) { offs.push_back(total); total += ((in.second * 16 + 255) / 256) * 256; }
  const size_t out_off = total;
  total += out_count * 16 + 256;
  RET(ensure_scratch(h, total));
  std::vector<double2*> dptr;
  for (size_t i = 0; i < ins.size(); ++i) {
    double2* d = reinterpret_cast<double2*>(h->d_scratch + offs[i]);
    if (ins[i].first) CU(h, cudaMemcpyAsync(d, ins[i].first, ins[i].second * 16, cudaMemcpyHostToDevice, h->stream));
    dptr.push_back(ins[i].first ? d : nullptr);
  }
  double2* dout = reinterpret_cast<double2*>(h->d_scratch + out_off);
  CU(h, body(dptr, dout));
  CU(h, cudaMemcpyAsync(out, dout, out_count * 16, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return QCB_OK;
}

int32_t qcb_la_matmul(qcb_handle h, const double* A, const double* B, uint64_t m, uint64_t k, uint64_t n, double* C) {
  if (h && h->is_group) return qcb_la_matmul(h->members[0], A, B, m, k, n, C);
  ENTER(h);
  if (!A || !B || !C) return fail(h, QCB_ERR_INVALID, "null argument");
  return la_run(h, {{A, m * k}, {B, k * n}}, m * n, C, [&](std::vector<double2*>& d, double2* o) { return launch_la_matmul(d[0], d[1], m, k, n, o, h->stream); });
}
int32_t qcb_la_matvec(qcb_handle h, const double* A, const double* x, uint64_t rows, uint64_t cols, double* y) {
  return qcb_la_matmul(h, A, x, rows, cols, 1, y);
}
int32_t qcb_la_kron(qcb_handle h, const double* A, uint64_t ar, uint64_t ac, const double* B, uint64_t br, uint64_t bc, double* C) {
  if (h && h->is_group) return qcb_la_kron(h->members[0], A, ar, ac, B, br, bc, C);
  ENTER(h);
  if (!A || !B || !C) return fail(h, QCB_ERR_INVALID, "null argument");
  return la_run(h, {{A, ar * ac}, {B, br * bc}}, ar * ac * br * bc, C, [&](std::vector<double2*>& d, double2* o) { return launch_la_kron(d[0], ar, ac, d[1], br, bc, o, h->stream); });
}
int32_t qcb_la_outer(qcb_handle h, const double* x, const double* y, uint64_t n, uint64_t m, double* C) {
  if (h && h->is_group) return qcb_la_outer(h->members[0], x, y, n, m, C);
  ENTER(h);
  if (!x || !y || !C) return fail(h, QCB_ERR_INVALID, "null argument");
  return la_run(h, {{x, n}, {y, m}}, n * m, C, [&](std::vector<double2*>& d, double2* o) { return launch_la_outer(d[0], d[1], n, m, o, h->stream); });
}
int32_t qcb_la_axpby(qcb_handle h, const double alpha[2], const double* x, const double beta[2], const double* y, uint64_t n, double* out) {
  if (h && h->is_group) return qcb_la_axpby(h->members[0], alpha, x, beta, y, n, out);
  ENTER(h);
  if (!alpha || !x || !out || (y && !beta)) return fail(h, QCB_ERR_INVALID, "null argument");
  const double2 al{alpha[0], alpha[1]}, be{beta ? beta[0] : 0.0, beta ? beta[1] : 0.0};
  return la_run(h, {{x, n}, {y, n}}, n, out, [&](std::vector<double2*>& d, double2* o) { return launch_la_axpby(al, d[0], be, d[1], n, o, h->stream); });
}
int32_t qcb_la_inner(qcb_handle h, const double* x, const double* y, uint64_t n, double out[2]) {
  if (h && h->is_group) return qcb_la_inner(h->members[0], x, y, n, out);
  ENTER(h);
  if (!x || !y || !out) return fail(h, QCB_ERR_INVALID, "null argument");
  const int grid = (int)std::min<uint64_t>((n + RED_THREADS - 1) / RED_THREADS, (uint64_t)red_grid(h));
  RET(ensure_partials(h, (size_t)grid * 2));
  RET(ensure_scratch(h, 2 * (n * 16 + 256)));
  double2* dx = reinterpret_cast<double2*>(h->d_scratch);
  double2* dy = reinterpret_cast<double2*>(h->d_scratch + ((n * 16 + 255) / 256) * 256);
  CU(h, cudaMemcpyAsync(dx, x, n * 16, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(dy, y, n * 16, cudaMemcpyHostToDevice, h->stream));
  CU(h, launch_inner(dx, dy, n, h->d_partials, grid, h->stream));     // conjugates the first argument
  CU(h, launch_finalize(h->d_partials, grid, 2, 0, 0.0, h->d_vals + 8, h->stream));
  return read_back(h, h->d_vals + 8, 16, out);
}
int32_t qcb_la_norm2(qcb_handle h, const double* x, uint64_t n, double* out) {
  double v[2];
  int rc = qcb_la_inner(h, x, x, n, v);
  if (rc != QCB_OK) return rc;
  *out = std::sqrt(v[0]);
  return QCB_OK;
}

// ---- P2, host side: decompositions, matrix functions, predicates (la_host.cpp).  Small dense matrices, no GPU work:
// `h` may be NULL (errors then go to the global slot read by qcb_last_error(NULL, ...)).
#define LA_ARGS(cond) do { if (!(cond)) return fail(h, QCB_ERR_INVALID, "null argument or empty matrix"); } while (0)
#define LA_LOCK std::unique_lock<std::recursive_mutex> _lk; if (h) _lk = std::unique_lock<std::recursive_mutex>(h->mu)
int32_t qcb_la_hadamard(qcb_handle h, const double* A, const double* B, uint64_t n, double* C) {
  LA_LOCK; LA_ARGS(A && B && C);
  la_hadamard(A, B, n, C); return QCB_OK;
}
int32_t qcb_la_transpose(qcb_handle h, const double* A, uint64_t rows, uint64_t cols, int32_t conjugate, double* out) {
  LA_LOCK; LA_ARGS(A && out);
  la_transpose(A, rows, cols, conjugate, out); return QCB_OK;
}
int32_t qcb_la_solve(qcb_handle h, const double* A, const double* B, uint64_t n, uint64_t nrhs, double* X) {
  LA_LOCK; LA_ARGS(A && B && X && n && nrhs);
  std::string err;
  return la_solve(A, B, n, nrhs, X, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_inverse(qcb_handle h, const double* A, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && n);
  std::string err;
  return la_inverse(A, n, out, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_is_hermitian(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out) {
  LA_LOCK; LA_ARGS(A && out);
  *out = la_is_hermitian(A, n, eps); return QCB_OK;
}
int32_t qcb_la_is_diagonal(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out) {
  LA_LOCK; LA_ARGS(A && out);
  *out = la_is_diagonal(A, n, eps); return QCB_OK;
}
int32_t qcb_la_is_unitary(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out) {
  LA_LOCK; LA_ARGS(A && out);
  *out = la_is_unitary(A, n, eps); return QCB_OK;
}
int32_t qcb_la_is_positive_semidefinite(qcb_handle h, const double* A, uint64_t n, double eps, int32_t* out) {
  LA_LOCK; LA_ARGS(A && out && n);
  std::string err;
  const int r = la_is_psd(A, n, eps, err);
  if (r < 0) return fail(h, QCB_ERR_INVALID, err);
  *out = r; return QCB_OK;
}
int32_t qcb_la_eigen_hermitian(qcb_handle h, const double* A, uint64_t n, double* eigenvalues, double* eigenvectors) {
  LA_LOCK; LA_ARGS(A && eigenvalues && eigenvectors && n);
  la_eigh(A, n, eigenvalues, eigenvectors); return QCB_OK;
}
int32_t qcb_la_eigen_general(qcb_handle h, const double* A, uint64_t n, double* eigenvalues, double* eigenvectors) {
  LA_LOCK; LA_ARGS(A && eigenvalues && eigenvectors && n);
  std::string err;
  return la_eig(A, n, eigenvalues, eigenvectors, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_svd(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* U, double* S, double* Vh) {
  LA_LOCK; LA_ARGS(A && U && S && Vh && m && n);
  la_svd(A, m, n, U, S, Vh); return QCB_OK;
}
int32_t qcb_la_lu(qcb_handle h, const double* A, uint64_t n, double* P, double* L, double* U) {
  LA_LOCK; LA_ARGS(A && P && L && U && n);
  la_lu(A, n, P, L, U); return QCB_OK;
}
int32_t qcb_la_qr(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* Q, double* R) {
  LA_LOCK; LA_ARGS(A && Q && R && m && n);
  la_qr(A, m, n, Q, R); return QCB_OK;
}
int32_t qcb_la_cholesky(qcb_handle h, const double* A, uint64_t n, double* L) {
  LA_LOCK; LA_ARGS(A && L && n);
  std::string err;
  return la_cholesky(A, n, L, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_matrix_exp(qcb_handle h, const double* A, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && n);
  la_expm(A, n, out); return QCB_OK;
}
int32_t qcb_la_matrix_log(qcb_handle h, const double* A, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && n);
  std::string err;
  return la_logm(A, n, out, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_matrix_sqrt(qcb_handle h, const double* A, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && n);
  std::string err;
  return la_sqrtm(A, n, out, err) == 0 ? QCB_OK : fail(h, QCB_ERR_STATE, err);
}
int32_t qcb_la_spectral_norm(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && m && n);
  std::vector<double> s;
  la_singular_values(A, m, n, s);
  *out = s.empty() ? 0.0 : s[0]; return QCB_OK;
}
int32_t qcb_la_condition_number(qcb_handle h, const double* A, uint64_t m, uint64_t n, double* out) {
  LA_LOCK; LA_ARGS(A && out && m && n);
  std::vector<double> s;
  la_singular_values(A, m, n, s);
  *out = (s.empty() || s.back() <= 0.0) ? HUGE_VAL : s[0] / s.back(); return QCB_OK;
}
#undef LA_ARGS
#undef LA_LOCK
int32_t qcb_la_trace(qcb_handle h, const double* A, uint64_t n, double out[2]) {
  if (h && h->is_group) return qcb_la_trace(h->members[0], A, n, out);
  ENTER(h);
  if (!A || !out) return fail(h, QCB_ERR_INVALID, "null argument");
  RET(ensure_scratch(h, n * n * 16 + 256));
  double2* dA = reinterpret_cast<double2*>(h->d_scratch);
  CU(h, cudaMemcpyAsync(dA, A, n * n * 16, cudaMemcpyHostToDevice, h->stream));
  CU(h, launch_la_trace(dA, n, h->d_vals + 8, h->stream));
  return read_back(h, h->d_vals + 8, 16, out);
}

}  // extern "C"

// =================================================================== job execution
namespace {

int run_job(qcb_sim* h, Job& job) {
  if (job.cancel.load()) { job.status.store(QCB_JOB_CANCELLED); return QCB_OK; }
  job.status.store(QCB_JOB_RUNNING);
  auto t0 = std::chrono::steady_clock::now();
  h->active_cancel = &job.cancel;
  int rc = QCB_OK;
  auto step = [&](int r) { if (rc == QCB_OK) rc = r; };
  if (!job.initial.empty()) step(qcb_set_state(h, job.initial.data(), job.initial.size() / 2));
  else step(qcb_set_zero(h));
  if (rc == QCB_OK) step(qcb_apply_ops(h, job.ops.data(), job.ops.size()));
  if (rc == QCB_OK && !job.uniforms.empty()) {
    job.outcomes.resize(job.uniforms.size());
    step(qcb_sample(h, job.uniforms.data(), job.uniforms.size(), job.outcomes.data()));
  }
  if (rc == QCB_OK && !job.ham_coeffs.empty()) {
    std::vector<const char*> ptrs;
    for (auto& s : job.ham_strings) ptrs.push_back(s.c_str());
    step(qcb_expect_hamiltonian(h, job.ham_coeffs.data(), ptrs.data(), ptrs.size(), &job.energy, nullptr));
    job.has_energy = (rc == QCB_OK);
  }
  if (rc == QCB_OK && job.want_probs) { job.probs.resize(h->local_count); step(qcb_probabilities(h, 0, h->local_count, job.probs.data())); }
  if (rc == QCB_OK && job.want_state) { job.state.resize(2 * h->local_count); step(qcb_get_state(h, 0, h->local_count, job.state.data())); }
  if (rc == QCB_OK) step(qcb_synchronize(h));
  h->active_cancel = nullptr;
  job.exec_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  // the inputs are not needed any more; the outputs of all but the 64 most recently finished jobs are dropped as well, so
  // that a variational loop through the job API does not grow the host heap (status and timing stay queryable)
  std::vector<qcb_op>().swap(job.ops); job.ext_d.clear(); job.ext_i.clear();
  std::vector<double>().swap(job.initial); std::vector<double>().swap(job.uniforms);
  if (job.cancel.load()) job.status.store(QCB_JOB_CANCELLED);
  else if (rc != QCB_OK) { job.error = h->err; job.status.store(QCB_JOB_FAILED); }   // never throw out of the worker (ideal_simulator.clj:93-96)
  else job.status.store(QCB_JOB_COMPLETED);
  {
    std::unique_lock<std::mutex> lk(h->jmu);
    h->finished.push_back(job.id);
    while (h->finished.size() > 64) {
      auto it = h->jobs.find(h->finished.front());
      h->finished.pop_front();
      if (it == h->jobs.end()) continue;
      Job& old = *it->second;
      std::vector<double>().swap(old.probs); std::vector<double>().swap(old.state); std::vector<uint64_t>().swap(old.outcomes);
      old.payload_dropped = true;
    }
  }
  return rc;
}

void worker_main(qcb_sim* h) {
  for (;;) {
    std::shared_ptr<Job> job;
    {
      std::unique_lock<std::mutex> lk(h->jmu);
      h->jcv.wait(lk, [&] { return h->stopping || !h->queue.empty(); });
      if (h->stopping && h->queue.empty()) return;
      job = h->queue.front();
      h->queue.pop_front();
    }
    if (job->status.load() == QCB_JOB_CANCELLED) continue;
    std::lock_guard<std::recursive_mutex> lk(h->mu);
    run_job(h, *job);
  }
}

}  // namespace
