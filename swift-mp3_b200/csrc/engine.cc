// Host engine + C ABI (include/mp3b200.h) of the B200-native MP3 encode path.
// SRC = Sources/SwiftMP3/MP3Encoder.swift of the reference.  There is no CPU fallback anywhere in this file: every
// encode call runs the CUDA pipeline of kernels.cu, and creation fails when no sm_100 device is usable.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mp3b200.h"
#include "kernels.h"
#include "tables.h"
#include "tables_gen.h"

using namespace mp3b;

namespace {

thread_local std::string g_err;
int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(e_ == cudaErrorMemoryAllocation ? MP3B_ERR_OOM : MP3B_ERR_CUDA, \
                                       "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <class T> inline T round_up(T v, T a) { return (v + a - 1) / a * a; }
constexpr int kMaxStreams = 65535;      // streams of one device batch: the stream index is gridDim.y in several launches

std::mutex g_dev_mu;
bool g_tables_uploaded[64] = {};

int usable_device(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p.major == 10;
}

}  // namespace

// One host thread per device of a multi-device batch (streams are independent, so a device batch
// never waits for another one).  The thread lives as long as the batch: it issues all CUDA work of its device and keeps the
// device current, and the calling thread only posts a job to every worker and waits for all of them.
struct PartWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, stop = false;
  int rc = 0;
  std::string err;
};

struct mp3b_batch {
  // ---- multi-device plane: a parent owns one device batch ("part") per entry of its device list and routes by stream index;
  // every field below this block is unused in a parent except S, opt and cfg
  std::vector<mp3b_batch *> parts;
  std::vector<int> part_lo;                              // first stream of part k; part_lo[parts.size()] = S
  std::vector<std::unique_ptr<PartWorker>> workers;
  Config cfg{};
  mp3b_options opt{};
  int device = 0, S = 0, ch = 0, Fc = 0, GC = 0;
  cudaStream_t st = nullptr;
  PassBuffers pb{};
  float *d_head[2] = {nullptr, nullptr};
  int head_sel = 0;
  float *d_stage[2] = {nullptr, nullptr}; size_t stage_stride = 0;   // double-buffered PCM staging, floats per stream
  int16_t *d_stage16[2] = {nullptr, nullptr};                        // the same for 16-bit input (converted into d_stage on the device)
  StreamPlan *h_plan = nullptr;                          // pinned [2][S]
  StreamPlan *d_plan[2] = {nullptr, nullptr};
  cudaStream_t st_copy = nullptr, st_d2h = nullptr;
  // The tail of a pass (scan, pack, frames: latency- and issue-bound, FMA pipe idle) runs on st_tail beside the head of the next
  // pass (filterbank, granule) on st.  What the head writes for the tail exists twice, selected by pass parity (alt = parity 1).
  cudaStream_t st_tail = nullptr;
  cudaEvent_t ev_head_done[2] = {}, ev_tail_done[2] = {};
  PassBuffers alt{};                                      // only the head -> tail arrays are set
  size_t h_pitch = 0;                                    // > 0: h_out is [S][h_pitch] (progressive download), else compact at h_offsets
  cudaEvent_t ev_h2d[2][2] = {}, ev_consumed[2] = {};
  StreamState *h_state = nullptr;                        // pinned [S]
  uint16_t *h_emit_size = nullptr; uint32_t *h_emit_n = nullptr;   // pinned [2][...]: one copy per pass in flight
  uint64_t *d_offsets = nullptr, *h_offsets = nullptr;
  uint8_t *d_compact = nullptr; size_t compact_cap = 0;
  uint8_t *h_out = nullptr; size_t h_out_cap = 0;
  size_t out_cap_bytes = 0;                              // allocation size of pb.out
  size_t out_total = 0;
  bool have_host_out = false;
  std::vector<uint32_t> pending;                         // floats waiting after the carried frame, per stream
  std::vector<uint8_t> fed;                              // the stream has been given samples (ISO mode level 3 flushes its delayed granule)
  std::vector<uint32_t> out_len;
  std::vector<uint32_t> frame_count, byte_count;
  std::vector<std::vector<uint16_t>> frame_sizes;       // SRC:258
  int max_frame_bytes = 0;
  int sticky = 0;
  bool planned = false;                                  // the current call has started to mutate the host bookkeeping
  // measurement
  float stage_ms[MP3B_STAGE_COUNT] = {};
  int launches = 0, passes = 0;
  cudaEvent_t ev[MP3B_STAGE_COUNT + 1] = {};
  cudaEvent_t evp[2][10] = {};                           // per-pass stage boundaries ([8] = the pass's results are on the host, [9] = end of the head)
  // trace
  int trace = 0;
  std::vector<std::vector<mp3b_frame_record>> tr_frames;
  std::vector<std::vector<mp3b_gc_record>> tr_gc;
  std::vector<std::vector<float>> tr_spec, tr_thr, tr_psy;
  std::vector<std::vector<int32_t>> tr_ix, tr_sf;
  PsyTab *d_psy = nullptr;                               // ISO mode level 2
  float *d_tc_b = nullptr;                               // tensor-core matrixing: split + swizzled analysis matrix
  std::vector<FrameRec> h_rec;
};

struct mp3b_session {
  mp3b_batch *b = nullptr;
  std::vector<uint8_t> pending_out;
};

namespace {

inline bool is_multi(const mp3b_batch *b) { return b && !b->parts.empty(); }
// part that owns `stream` of a parent, *local = its index inside the part
inline mp3b_batch *part_of(const mp3b_batch *b, int stream, int *local) {
  int k = (int)(std::upper_bound(b->part_lo.begin(), b->part_lo.end(), stream) - b->part_lo.begin()) - 1;
  k = std::min(std::max(k, 0), (int)b->parts.size() - 1);
  *local = stream - b->part_lo[k];
  return b->parts[k];
}
void part_worker_main(PartWorker *w) {
  std::unique_lock<std::mutex> lk(w->m);
  for (;;) {
    w->cv.wait(lk, [&] { return w->stop || w->has_job; });
    if (w->stop) return;
    std::function<int()> job = std::move(w->job);
    w->has_job = false;
    lk.unlock();
    const int rc = job();
    std::string err = rc ? g_err : std::string();
    lk.lock();
    w->rc = rc; w->err = std::move(err); w->done = true;
    w->cv.notify_all();
  }
}
// fn(k) runs on worker k for every part at once; returns the first failure (its text becomes the caller's last error)
int on_all_parts(const std::vector<std::unique_ptr<PartWorker>> &workers, const std::function<int(int)> &fn) {
  for (size_t k = 0; k < workers.size(); ++k) {
    PartWorker &w = *workers[k];
    std::lock_guard<std::mutex> lk(w.m);
    w.job = [&fn, k] { return fn((int)k); };
    w.has_job = true; w.done = false;
    w.cv.notify_all();
  }
  int rc = MP3B_OK;
  for (auto &wp : workers) {
    PartWorker &w = *wp;
    std::unique_lock<std::mutex> lk(w.m);
    w.cv.wait(lk, [&] { return w.done; });
    if (w.rc && !rc) { rc = w.rc; g_err = w.err; }
  }
  return rc;
}
void stop_workers(std::vector<std::unique_ptr<PartWorker>> &workers) {
  for (auto &wp : workers) {
    { std::lock_guard<std::mutex> lk(wp->m); wp->stop = true; }
    wp->cv.notify_all();
    if (wp->th.joinable()) wp->th.join();
  }
  workers.clear();
}

int fill_config(const mp3b_options &o, int n_streams, Config &c) {
  if (o.sample_rate <= 0) return fail(MP3B_ERR_BAD_ARG, "sample_rate must be > 0");
  if (o.mode < 0 || o.mode > 2) return fail(MP3B_ERR_BAD_ARG, "mode must be 0 (mono), 1 (stereo) or 2 (jointStereo)");
  memset(&c, 0, sizeof c);
  c.n_streams = n_streams;
  c.channels = o.mode == 0 ? 1 : 2;                              // SRC:300
  c.fsc = 1152 * c.channels;
  c.sample_rate = o.sample_rate; c.base_kbps = o.bitrate_kbps; c.vbr = o.vbr ? 1 : 0; c.mode = o.mode;
  c.quality = std::min(9, std::max(0, o.quality));               // SRC:110
  c.crc = o.crc_protected ? 1 : 0; c.original = o.original ? 1 : 0; c.copyright = o.copyright ? 1 : 0;
  c.sr_index = sample_rate_index(o.sample_rate); c.sfb_index = sfb_table_index(o.sample_rate);
  c.side_bytes = c.channels == 1 ? 17 : 32;
  c.header_bytes = 4 + (c.crc ? 2 : 0) + c.side_bytes;
  if (o.mode == 0) { c.mode_bits = 3; c.mode_ext = 0; } else if (o.mode == 2) { c.mode_bits = 1; c.mode_ext = 2; } else { c.mode_bits = 0; c.mode_ext = 0; }
  c.cbr_index = bitrate_index(o.bitrate_kbps, o.sample_rate);
  c.f_one = 1.0f; c.f_neg0 = -0.0f;
  c.iso = 0; c.ms_scale = 0.5f;                                  // SRC:2148-2154; mp3b_batch_set_iso_mode changes both
  for (int i = 0; i < 16; ++i) {
    long long num = 144LL * bitrate_value(i) * 1000;             // SRC:490-495
    c.frame_base[i] = (int)(num / o.sample_rate); c.frame_rem[i] = (int)(num % o.sample_rate);
  }
  for (int k = 0; k <= 320; ++k) c.vbr_idx_of_kbps[k] = (uint8_t)bitrate_index(k, o.sample_rate);
  // every bitrate index the session can reach must leave room for header + side info (the reference traps otherwise)
  int lo_idx = c.cbr_index, hi_idx = c.cbr_index;
  if (c.vbr) {
    int lo = std::max(32, c.base_kbps - 64 + c.quality * 8), hi = std::min(320, c.base_kbps + 64 - c.quality * 4);
    lo_idx = 15; hi_idx = 0;
    for (int k = std::min(lo, hi); k <= std::max(lo, hi); ++k) {     // lo may exceed 320 (base 320, quality 9): k_bitrate clamps the same way
      const int kk = std::min(std::max(k, 0), 320);
      lo_idx = std::min<int>(lo_idx, c.vbr_idx_of_kbps[kk]); hi_idx = std::max<int>(hi_idx, c.vbr_idx_of_kbps[kk]);
    }
  }
  for (int i = lo_idx; i <= hi_idx; ++i)
    if (c.frame_base[i] - c.header_bytes < 2 * c.channels || c.frame_base[i] + 1 > 65000)
      return fail(MP3B_ERR_BAD_ARG, "bitrate index %d at %d Hz gives a %d-byte frame: unusable", i, o.sample_rate, c.frame_base[i]);
  return MP3B_OK;
}

int max_frame_bytes_of(const Config &c) {
  int m = 0;
  if (!c.vbr) return c.frame_base[c.cbr_index] + 1;
  for (int k = 0; k <= 320; ++k) m = std::max(m, c.frame_base[c.vbr_idx_of_kbps[k]] + 1);
  return m;
}

void free_batch(mp3b_batch *b) {
  if (!b) return;
  if (is_multi(b) || !b->workers.empty()) {
    stop_workers(b->workers);
    for (mp3b_batch *p : b->parts) free_batch(p);
    delete b;
    return;
  }
  cudaSetDevice(b->device);
  if (b->st) cudaStreamSynchronize(b->st);
  PassBuffers &p = b->pb;
  void *dev[] = {p.plan, p.state, b->d_head[0], b->d_head[1], p.ms, p.frame_energy, p.gc_energy, p.gc_bt, p.frame_br, p.spec, p.sub, p.smag,
                 p.gc_meta, p.gc_bits, p.gc_bv, p.gc_bitoff, p.gc_sel, p.fr_md, p.rec, p.emit, p.md, p.md_tail, p.md_carry, p.out,
                 p.emit_size, p.emit_n, p.tr_ix, p.tr_thr, p.gc_psy, p.gc_sf, b->d_psy, b->d_tc_b, b->alt.smag, b->alt.gc_meta, b->alt.gc_bits, b->alt.gc_bv, b->alt.gc_bt, b->alt.gc_energy, b->alt.frame_br,
                 b->alt.frame_energy, b->alt.ms, b->alt.gc_psy, b->alt.gc_sf, b->d_stage[0], b->d_stage[1], b->d_stage16[0], b->d_stage16[1], b->d_plan[1], b->d_offsets, b->d_compact};
  for (void *q : dev) if (q) cudaFree(q);
  void *host[] = {b->h_plan, b->h_state, b->h_emit_size, b->h_emit_n, b->h_offsets, b->h_out};
  for (void *q : host) if (q) cudaFreeHost(q);
  for (auto &e : b->ev) if (e) cudaEventDestroy(e);
  for (auto &e : b->ev_consumed) if (e) cudaEventDestroy(e);
  for (auto &e : b->ev_head_done) if (e) cudaEventDestroy(e);
  for (auto &e : b->ev_tail_done) if (e) cudaEventDestroy(e);
  if (b->st_tail) { cudaStreamSynchronize(b->st_tail); cudaStreamDestroy(b->st_tail); }
  for (auto &r : b->evp) for (auto &e : r) if (e) cudaEventDestroy(e);
  for (auto &r : b->ev_h2d) for (auto &e : r) if (e) cudaEventDestroy(e);
  if (b->st) cudaStreamDestroy(b->st);
  if (b->st_copy) cudaStreamDestroy(b->st_copy);
  if (b->st_d2h) cudaStreamDestroy(b->st_d2h);
  cudaGetLastError();
  delete b;
}

template <class T> cudaError_t dalloc(T *&p, size_t n, bool zero = true) {
  cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
  if (e == cudaSuccess && zero) e = cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T));
  return e;
}

int create_batch(const mp3b_options *opts, int n_streams, int device, int frames_per_pass, mp3b_batch **out) {
  if (!opts || !out || n_streams <= 0) return fail(MP3B_ERR_BAD_ARG, "null options / out or n_streams <= 0");
  if (n_streams > kMaxStreams) return fail(MP3B_ERR_BAD_ARG, "n_streams %d exceeds the per-device limit of %d (the stream index is a grid y dimension); split the batch or use mp3b_batch_create_multi", n_streams, kMaxStreams);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return fail(MP3B_ERR_CUDA, "no CUDA device: the engine has no CPU fallback"); }
  if (device < 0 || device >= count) return fail(MP3B_ERR_BAD_ARG, "device %d out of range (%d devices)", device, count);
  if (!usable_device(device)) return fail(MP3B_ERR_CUDA, "device %d is not compute capability 10.x (sm_100a kernels only)", device);
  Config cfg;
  int rc = fill_config(*opts, n_streams, cfg);
  if (rc) return rc;
  CU(cudaSetDevice(device));
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (device < 64 && !g_tables_uploaded[device]) { CU(upload_tables()); g_tables_uploaded[device] = true; }
  }
  mp3b_batch *b = new mp3b_batch();
  b->cfg = cfg; b->opt = *opts; b->opt.quality = cfg.quality; b->device = device; b->S = n_streams; b->ch = cfg.channels;
  int Fc = frames_per_pass > 0 ? frames_per_pass : (int)std::min<long long>(4096, std::max<long long>(8, 262144 / n_streams));
  b->Fc = Fc; b->GC = Fc * 2 * cfg.channels;
  b->max_frame_bytes = max_frame_bytes_of(cfg);
  const size_t S = n_streams, GC = b->GC;
  PassBuffers &p = b->pb;
  p.Fc = Fc; p.GC = b->GC;
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&b->st_copy, cudaStreamNonBlocking));
  A(cudaStreamCreateWithFlags(&b->st_d2h, cudaStreamNonBlocking));
  {
    int lo_prio = 0, hi_prio = 0;
    cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);          // the tail's short kernels go first whenever an SM has room
    A(cudaStreamCreateWithPriority(&b->st_tail, cudaStreamNonBlocking, hi_prio));
  }
  for (auto &ev : b->ev_head_done) A(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : b->ev_tail_done) A(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  A(dalloc(p.plan, S)); A(dalloc(b->d_plan[1], S)); A(dalloc(p.state, S));
  A(dalloc(b->d_head[0], S * 2 * cfg.fsc)); A(dalloc(b->d_head[1], S * 2 * cfg.fsc));
  A(dalloc(p.ms, S * (Fc + 1))); A(dalloc(p.frame_energy, S * Fc)); A(dalloc(p.gc_energy, S * (10 + GC)));
  A(dalloc(p.gc_bt, S * GC)); A(dalloc(p.frame_br, S * Fc)); A(dalloc(p.smag, S * GC * 576, false));
  p.sub_rows = 18 * (1 + 2 * Fc);                                // + the carried granule in front (MDCT overlap), zero before the first frame
  A(dalloc(p.sub, S * cfg.channels * (size_t)p.sub_rows * 32, false));
  A(cudaMemset2D(p.sub, (size_t)p.sub_rows * 32 * sizeof(float), 0, 576 * sizeof(float), S * cfg.channels));
  A(dalloc(p.gc_meta, S * GC)); A(dalloc(p.gc_bits, S * GC * kMaxEntries)); A(dalloc(p.gc_bv, S * GC * kMaxEntries));
  if (getenv("MP3B_OVERLAP")) {                                // second copy of what a pass's head hands to its tail (two-stream pipeline, opt-in)
    PassBuffers &q = b->alt;
    A(dalloc(q.ms, S * (Fc + 1))); A(dalloc(q.frame_energy, S * Fc)); A(dalloc(q.gc_energy, S * (10 + GC)));
    A(dalloc(q.gc_bt, S * GC)); A(dalloc(q.frame_br, S * Fc)); A(dalloc(q.smag, S * GC * 576, false));
    A(dalloc(q.gc_meta, S * GC)); A(dalloc(q.gc_bits, S * GC * kMaxEntries)); A(dalloc(q.gc_bv, S * GC * kMaxEntries));
  }
  A(dalloc(p.gc_bitoff, S * GC)); A(dalloc(p.gc_sel, S * GC)); A(dalloc(p.fr_md, S * Fc * 2)); A(dalloc(p.rec, S * (Fc + 1))); A(dalloc(p.emit, S * (Fc + 1)));
  p.md_stride = round_up<size_t>(kMdCarryCap + (size_t)Fc * 2 * cfg.channels * 540, 16);
  A(dalloc(p.md, S * p.md_stride + 16, false));   // + one word: k_frames reads aligned word pairs
  A(dalloc(p.md_tail, S * 4)); A(dalloc(p.md_carry, S * kMdCarryCap));
  A(dalloc(p.emit_size, S * (Fc + 1))); A(dalloc(p.emit_n, S));
  A(dalloc(b->d_offsets, S + 1));
  A(cudaHostAlloc((void **)&b->h_plan, 2 * S * sizeof(StreamPlan), cudaHostAllocDefault));
  A(cudaHostAlloc((void **)&b->h_state, S * sizeof(StreamState), cudaHostAllocDefault));
  A(cudaHostAlloc((void **)&b->h_emit_size, 2 * S * (Fc + 1) * sizeof(uint16_t), cudaHostAllocDefault));
  A(cudaHostAlloc((void **)&b->h_emit_n, 2 * S * sizeof(uint32_t), cudaHostAllocDefault));
  A(cudaHostAlloc((void **)&b->h_offsets, (S + 1) * sizeof(uint64_t), cudaHostAllocDefault));
  for (auto &ev : b->ev) A(cudaEventCreate(&ev));
  for (auto &ev : b->ev_consumed) A(cudaEventCreate(&ev));
  for (auto &r : b->evp) for (auto &ev : r) A(cudaEventCreate(&ev));
  for (auto &r : b->ev_h2d) for (auto &ev : r) A(cudaEventCreate(&ev));
  if (e != cudaSuccess) {
    free_batch(b);
    return fail(e == cudaErrorMemoryAllocation ? MP3B_ERR_OOM : MP3B_ERR_CUDA, "batch allocation failed: %s", cudaGetErrorString(e));
  }
  if (!cfg.vbr && (cudaMemset(p.frame_br, cfg.cbr_index, S * Fc) != cudaSuccess || (b->alt.frame_br && cudaMemset(b->alt.frame_br, cfg.cbr_index, S * Fc) != cudaSuccess))) { free_batch(b); return fail(MP3B_ERR_CUDA, "batch initialisation failed"); }   // constant for CBR: the pre-pass may be skipped
  // The zero fills above ran on the legacy default stream; the engine's own streams are non-blocking and never synchronise
  // with it, so the batch is handed out only when they have landed.
  if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) { free_batch(b); return fail(MP3B_ERR_CUDA, "batch initialisation failed"); }
  b->pending.assign(S, 0); b->out_len.assign(S, 0); b->frame_count.assign(S, 0); b->byte_count.assign(S, 0);
  b->fed.assign(S, 0);
  b->frame_sizes.resize(S);
  b->d_plan[0] = p.plan;
  // A small batch (a session is a batch of one) gets its staging, output and pinned download buffers now instead of inside its
  // first encode call: four allocations, one of them pinned, are most of what a single short stream costs end to end.
  const size_t stage_bytes = S * (size_t)Fc * cfg.fsc * sizeof(float);
  if (stage_bytes <= ((size_t)64 << 20)) {
    b->stage_stride = (size_t)Fc * cfg.fsc;
    // (output for calls of up to four passes: growing the pinned download buffer inside a call costs more than the call)
    const size_t out_stride = round_up<size_t>((size_t)(4 * Fc + 2) * b->max_frame_bytes, 16), total = S * out_stride;
    bool ok = cudaMalloc((void **)&b->d_stage[0], stage_bytes) == cudaSuccess && cudaMalloc((void **)&b->d_stage[1], stage_bytes) == cudaSuccess &&
              cudaMalloc((void **)&b->pb.out, total) == cudaSuccess;
    if (ok) { b->out_cap_bytes = total; b->pb.out_stride = out_stride; }
    const size_t cap = round_up<size_t>(total + total / 8 + 4096, 4096);
    if (ok && cudaMalloc((void **)&b->d_compact, cap) == cudaSuccess) b->compact_cap = cap; else ok = false;
    if (ok && cudaHostAlloc((void **)&b->h_out, cap, cudaHostAllocDefault) == cudaSuccess) b->h_out_cap = cap; else ok = false;
    if (!ok) { cudaGetLastError(); free_batch(b); return fail(MP3B_ERR_OOM, "batch allocation failed"); }
  }
  *out = b;
  return MP3B_OK;
}

// Tables of the psychoacoustic model (iso_psy.cuh; restated by tests/psymodel.py).  ISO 11172-3 Annex D gives its partition tables
// as literals per sample rate; here they are derived from the same ingredients: a Bark scale (Zwicker / Terhardt), partitions of
// 1/3 Bark, the ISO model-2 spreading function, a minimum SNR that falls from 24.5 dB at the lowest partitions to 4.5 dB, and the
// absolute threshold of hearing (Terhardt) with a full-scale sine at 96 dB SPL.
double bark_of(double f) { return 13.0 * atan(0.00076 * f) + 3.5 * atan((f / 7500.0) * (f / 7500.0)); }
void build_psy_tab(int sample_rate, int sfb_index, PsyTab &t) {
  memset(&t, 0, sizeof t);
  const double fs = sample_rate, pi = 3.14159265358979323846;
  int np = 0, lo = 0;
  while (lo < 512 && np < kPsyMaxPart) {
    int hi = lo + 1;
    while (hi < 512 && bark_of(hi * fs / 1024.0) - bark_of(lo * fs / 1024.0) < 1.0 / 3.0) ++hi;
    if (np == kPsyMaxPart - 1) hi = 512;
    t.part_lo[np] = (uint16_t)lo; t.part_n[np] = (uint16_t)(hi - lo);
    for (int k = lo; k < hi; ++k) t.line_part[k] = (uint8_t)np;
    lo = hi; ++np;
  }
  t.n_part = np;
  double bval[kPsyMaxPart];
  for (int b = 0; b < np; ++b) bval[b] = bark_of((t.part_lo[b] + 0.5 * (t.part_n[b] - 1)) * fs / 1024.0);
  for (int i = 0; i < np; ++i) {                       // i = target, j = source
    double sum = 0.0;
    for (int j = 0; j < np; ++j) {
      double tx = (j >= i ? 3.0 : 1.5) * (bval[i] - bval[j]);
      double x = 0.0;
      if (tx >= 0.5 && tx <= 2.5) { const double u = tx - 0.5; x = 8.0 * (u * u - 2.0 * u); }
      tx += 0.474;
      const double ty = 15.811389 + 7.5 * tx - 17.5 * sqrt(1.0 + tx * tx);
      const double v = ty <= -60.0 ? 0.0 : pow(10.0, (x + ty) / 10.0);
      t.s3t[j * kPsyMaxPart + i] = (float)v;
      sum += v;
    }
    t.rnorm[i] = (float)(1.0 / sum);
    t.minval[i] = (float)std::min(24.5, std::max(4.5, 24.5 - 2.0 * bval[i]));
    double q = 0.0;
    for (int k = t.part_lo[i]; k < t.part_lo[i] + t.part_n[i]; ++k) {
      const double f = std::max(k * fs / 1024.0, 20.0) / 1000.0;
      const double ath = std::min(3.64 * pow(f, -0.8) - 6.5 * exp(-0.6 * (f - 3.3) * (f - 3.3)) + 1e-3 * f * f * f * f, 80.0);
      q += (32768.0 * 256.0) * (32768.0 * 256.0) * pow(10.0, (ath - 96.0) / 10.0);
    }
    t.qthr[i] = (float)q;
  }
  const int *cum = host_sfb_cum() + 21 * sfb_index;
  t.sfb_line[0] = 0;
  for (int i = 0; i < 21; ++i) t.sfb_line[i + 1] = (cum[i] * 8 + 4) / 9;
  t.sfb_line[22] = 512;
  for (int n = 0; n < 1024; ++n) t.hann1024[n] = (float)(0.5 * (1.0 - cos(2.0 * pi * (n + 0.5) / 1024.0)));
  for (int n = 0; n < 256; ++n) t.hann256[n] = (float)(0.5 * (1.0 - cos(2.0 * pi * (n + 0.5) / 256.0)));
  for (int j = 0; j < 768; ++j) { t.tw[j].x = (float)cos(2.0 * pi * j / 1024.0); t.tw[j].y = (float)-sin(2.0 * pi * j / 1024.0); }
}
int ensure_iso2(mp3b_batch *b) {
  PassBuffers &p = b->pb;
  if (p.gc_psy) return MP3B_OK;
  CU(cudaSetDevice(b->device));
  const size_t n = (size_t)b->S * b->GC * 24;
  CU(dalloc(p.gc_psy, n)); CU(dalloc(p.gc_sf, n));
  if (b->alt.smag) { CU(dalloc(b->alt.gc_psy, n)); CU(dalloc(b->alt.gc_sf, n)); }
  std::unique_ptr<PsyTab> t(new PsyTab);
  build_psy_tab(b->cfg.sample_rate, b->cfg.sfb_index, *t);
  CU(cudaMalloc((void **)&b->d_psy, sizeof(PsyTab)));
  CU(cudaMemcpy(b->d_psy, t.get(), sizeof(PsyTab), cudaMemcpyHostToDevice));
  p.psy = b->d_psy;
  CU(cudaStreamSynchronize(cudaStreamLegacy));
  return MP3B_OK;
}

int ensure_out(mp3b_batch *b, size_t stride) {
  stride = round_up<size_t>(stride, 16);
  size_t need = stride * b->S;
  if (need > b->out_cap_bytes) {
    if (b->pb.out) cudaFree(b->pb.out);
    b->pb.out = nullptr; b->out_cap_bytes = 0;
    CU(cudaMalloc((void **)&b->pb.out, need));
    b->out_cap_bytes = need;
  }
  b->pb.out_stride = stride;
  return MP3B_OK;
}

int ensure_trace(mp3b_batch *b) {
  PassBuffers &p = b->pb;
  const size_t n = (size_t)b->S * b->GC * 576;
  if ((b->trace & 5) && !p.spec) CU(dalloc(p.spec, n, false));      // spectrum trace / input of the threshold trace
  if ((b->trace & 2) && !p.tr_ix) CU(dalloc(p.tr_ix, n));
  if ((b->trace & 4) && !p.tr_thr) CU(dalloc(p.tr_thr, n));
  return MP3B_OK;
}

// One API call = encode(samples:) on every stream (+ optional flush()), split into passes of at most Fc frames.
int run_call_impl(mp3b_batch *b, const float *const *pcm, const size_t *n_floats, bool device_ptrs, int flush,
                  const uint8_t *flush_mask, bool download, size_t row_floats, int elem_bytes) {
  // elem_bytes = 2: pcm[] really are const int16_t * (mp3b_batch_encode_i16); sample i means Float(pcm[i]) / 32768
  if (!b) return fail(MP3B_ERR_BAD_ARG, "null batch");
  if (b->sticky) return fail(b->sticky, "batch is in a failed state: %s", g_err.c_str());
  CU(cudaSetDevice(b->device));
  const Config &cfg = b->cfg;
  const int S = b->S, fsc = cfg.fsc, Fc = b->Fc;
  std::vector<size_t> n(S, 0), cursor(S, 0);
  std::vector<uint8_t> want_flush(S, 0), flushed(S, 0);
  size_t max_frames = 0;
  for (int s = 0; s < S; ++s) {
    n[s] = (n_floats && pcm && pcm[s]) ? n_floats[s] : 0;
    want_flush[s] = flush && (!flush_mask || flush_mask[s]);
    size_t avail = b->pending[s] + n[s];
    max_frames = std::max(max_frames, avail / fsc + 1 + (cfg.iso >= 3 ? 1 : 0));
    if (n[s]) b->fed[s] = 1;
  }
  int rc = ensure_out(b, (max_frames + 1) * (size_t)b->max_frame_bytes);
  if (rc) return rc;
  // Host input is bound by PCIe, so the call is cut into about a dozen passes: the kernels and the download of pass p hide
  // behind the upload of pass p + 1 and only the last pass is exposed.  Device input keeps the largest passes.
  // ... as long as a pass still uploads some 16 MB: below that the fixed cost of a pass (8 launches, events, a host wake-up) is larger
  // than the copy it hides, so a single stream of a few seconds goes through in one or two passes (C1: 8 passes -> 1)
  const size_t min_pass_frames = ((size_t)16 << 20) / ((size_t)S * fsc * (size_t)elem_bytes) + 1;
  const int Fp = device_ptrs ? Fc : (int)std::min<size_t>((size_t)Fc, std::max<size_t>(std::max<size_t>((max_frames + 11) / 12, (size_t)std::min(Fc, 48)), min_pass_frames));
  // progressive download: after every pass the byte columns that pass produced are copied for all streams with one
  // strided D2H on its own stream into a pitched pinned buffer [S][out_stride]
  bool progressive = download && !device_ptrs && max_frames > (size_t)Fp;   // single-pass calls keep the compact copy
  std::vector<uint32_t> opos;
  if (progressive) {
    const size_t need = (size_t)S * b->pb.out_stride;
    if (need > b->h_out_cap) {
      if (b->h_out) cudaFreeHost(b->h_out);
      b->h_out = nullptr; b->h_out_cap = 0;
      CU(cudaHostAlloc((void **)&b->h_out, need, cudaHostAllocDefault));
      b->h_out_cap = need;
    }
    opos.assign(S, 0);
  }
  b->h_pitch = 0;
  if (b->trace) { rc = ensure_trace(b); if (rc) return rc; }
  if (!device_ptrs && !b->d_stage[0]) {
    b->stage_stride = (size_t)Fc * fsc;
    CU(cudaMalloc((void **)&b->d_stage[0], (size_t)S * b->stage_stride * sizeof(float)));
    CU(cudaMalloc((void **)&b->d_stage[1], (size_t)S * b->stage_stride * sizeof(float)));
  }
  if (elem_bytes == 2 && !b->d_stage16[0]) {
    CU(cudaMalloc((void **)&b->d_stage16[0], (size_t)S * b->stage_stride * sizeof(int16_t)));
    CU(cudaMalloc((void **)&b->d_stage16[1], (size_t)S * b->stage_stride * sizeof(int16_t)));
  }
  for (auto &m : b->stage_ms) m = 0.0f;
  b->launches = 0; b->passes = 0;
  b->have_host_out = false;
  if (b->trace) {
    b->tr_frames.assign(S, {}); b->tr_gc.assign(S, {}); b->tr_spec.assign(S, {}); b->tr_ix.assign(S, {}); b->tr_thr.assign(S, {});
    b->tr_psy.assign(S, {}); b->tr_sf.assign(S, {});
  }
  cudaStream_t st = b->st, stc = b->st_copy;
  // two-stream pass pipeline (see mp3b_batch::st_tail); the trace plane reads per-pass device buffers synchronously and keeps one stream
  // Opt-in (MP3B_OVERLAP=1 at batch creation) because it measured as a wash on the B200: 107.8 ms per C4 step with, 108.3 ms without
  // — k_filterbank (165 registers x 128 threads x 3 CTAs) and k_granule (64 x 256 x 4) fill the register file, so tail CTAs displace
  // head CTAs instead of sharing the SM with them.
  const bool overlap = !b->trace && b->alt.smag != nullptr;
  cudaStream_t stt = overlap ? b->st_tail : st;
  bool tail_used[2] = {false, false};
  // Pass pipeline: while the kernels of pass p run on `st`, the PCM of pass p + 1 crosses PCIe on `st_copy` into the
  // other staging buffer.  Planning is pure host arithmetic (it never needs device results), so it runs one pass ahead.
  std::vector<const float *> src[2] = {std::vector<const float *>(S), std::vector<const float *>(S)};
  bool first_plan = true;
  auto plan_pass = [&](int slot) -> bool {
    bool any = false;
    StreamPlan *hp = b->h_plan + (size_t)slot * S;
    for (int s = 0; s < S; ++s) {
      StreamPlan &pl = hp[s];
      size_t remaining = n[s] - cursor[s];
      size_t room = (size_t)Fp * fsc - b->pending[s];
      size_t cur_n = std::min(remaining, room);
      size_t total = b->pending[s] + cur_n;
      uint32_t nfr = (uint32_t)(total / fsc), flags = first_plan ? 4u : 0u;
      uint32_t new_pending = (uint32_t)(total % fsc);
      if (remaining == cur_n && want_flush[s] && !flushed[s]) {
        // ISO mode level 3 codes the signal 576 samples late: the last granule of the input still sits in the delay when the
        // padded frame ends before it does — one more frame (its PCM beyond the input reads as zeros) brings it out
        const uint32_t tail = cfg.iso >= 3 && b->fed[s] && (new_pending == 0 || new_pending > 576u * (uint32_t)cfg.channels) ? 1u : 0u;
        if (new_pending > 0 || tail) {
          const uint32_t add = (new_pending > 0 ? 1u : 0u) + tail;
          if (nfr + add <= (uint32_t)Fc) { nfr += add; flags |= 3u; new_pending = 0; flushed[s] = 1; b->fed[s] = 0; }   // (flush() stays idempotent)
        } else { flags |= 2u; flushed[s] = 1; }
      }
      src[slot][s] = (pcm && pcm[s]) ? (const float *)((const char *)pcm[s] + cursor[s] * (size_t)elem_bytes) : nullptr;
      pl.cur = device_ptrs ? src[slot][s] : b->d_stage[slot] + (size_t)s * b->stage_stride;
      pl.cur_n = (uint32_t)cur_n; pl.n_frames = nfr; pl.flags = flags; pl.head_n = (uint32_t)(fsc + b->pending[s]);
      if (cur_n || nfr || (flags & 2u)) any = true;
      cursor[s] += cur_n;
      b->pending[s] = new_pending;
    }
    const bool run = any || first_plan;
    first_plan = false;
    return run;
  };
  auto issue_h2d = [&](int slot) -> int {
    const StreamPlan *hp = b->h_plan + (size_t)slot * S;
    CU(cudaEventRecord(b->ev_h2d[slot][0], stc));
    if (!device_ptrs) {
      CU(cudaStreamWaitEvent(stc, b->ev_consumed[slot], 0));       // the kernels that last read this staging buffer
      // one strided copy when every stream that hands over data hands over the same amount and all buffers are equally
      // spaced (streams without data in this pass may sit in between: their rows are copied too and never read)
      size_t cur0 = 0;
      for (int s = 0; s < S; ++s) cur0 = std::max<size_t>(cur0, hp[s].cur_n);
      bool uniform = S > 1 && cur0 > 0;
      ptrdiff_t pitch = 0;
      for (int s = 0; s < S && uniform; ++s) {
        // a stream without data in this pass may only sit in the copy if its row is known to be readable that far
        const bool idle_ok = row_floats && src[slot][s] && (size_t)((const char *)src[slot][s] - (const char *)pcm[s]) / elem_bytes + cur0 <= row_floats;
        if (!src[slot][s] || (hp[s].cur_n != cur0 && !(hp[s].cur_n == 0 && idle_ok))) { uniform = false; break; }
        if (s >= 1) {
          ptrdiff_t d = (const char *)src[slot][s] - (const char *)src[slot][s - 1];
          if (s == 1) pitch = d; else if (d != pitch) uniform = false;
        }
      }
      char *stage = elem_bytes == 2 ? (char *)b->d_stage16[slot] : (char *)b->d_stage[slot];
      const size_t eb = (size_t)elem_bytes;
      if (uniform && pitch >= (ptrdiff_t)(cur0 * eb)) {
        CU(cudaMemcpy2DAsync(stage, b->stage_stride * eb, src[slot][0], (size_t)pitch, cur0 * eb, S, cudaMemcpyHostToDevice, stc));
      } else {
        for (int s = 0; s < S; ++s)
          if (hp[s].cur_n) CU(cudaMemcpyAsync(stage + (size_t)s * b->stage_stride * eb, src[slot][s], hp[s].cur_n * eb, cudaMemcpyHostToDevice, stc));
      }
    }
    if (tail_used[slot]) CU(cudaStreamWaitEvent(stc, b->ev_tail_done[slot], 0));   // the tail that last read this plan slot (slot == pass parity)
    CU(cudaMemcpyAsync(b->d_plan[slot], hp, (size_t)S * sizeof(StreamPlan), cudaMemcpyHostToDevice, stc));
    if (!device_ptrs && elem_bytes == 2) {                             // widen on the copy stream, behind the previous pass's kernels
      int k = launch_widen_i16(b->d_stage16[slot], b->d_stage[slot], b->stage_stride, b->d_plan[slot], S, stc);
      if (k < 0) return fail(MP3B_ERR_CUDA, "launch_widen_i16: %s", cudaGetErrorString((cudaError_t)(-k)));
      b->launches += k;
    }
    CU(cudaEventRecord(b->ev_h2d[slot][1], stc));
    return MP3B_OK;
  };
  // A pass's host-side work (frame sizes for the Xing TOC, the progressive download of its bytes, stage times) needs its
  // results on the host, but the next pass does not: the kernels of pass p + 1 are queued before the host waits for
  // pass p, so the device never idles between passes.  Only the trace plane (which reads per-pass device buffers with
  // synchronous copies) finishes every pass before the next one starts.
  auto finish_pass = [&](int slot, int par) -> int {
    const StreamPlan *hplan = b->h_plan + (size_t)slot * S;
    cudaError_t se = cudaEventSynchronize(b->evp[par][8]);
    if (se != cudaSuccess) { b->sticky = MP3B_ERR_CUDA; return fail(MP3B_ERR_CUDA, "device pipeline failed: %s", cudaGetErrorString(se)); }
    {
      static const int stage_of[7] = {MP3B_STAGE_H2D, MP3B_STAGE_PREPASS, MP3B_STAGE_SPECTRUM, MP3B_STAGE_CURVE, MP3B_STAGE_SCAN, MP3B_STAGE_PACK, MP3B_STAGE_FRAMES};
      for (int i = 1; i < 7; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, b->evp[par][i], b->evp[par][i == 3 ? 9 : i + 1]);   // the head ends at [9]; [4] is the start of the tail on its own stream
        b->stage_ms[stage_of[i]] += ms;
      }
      { float ms = 0; cudaEventElapsedTime(&ms, b->ev_h2d[slot][0], b->ev_h2d[slot][1]); b->stage_ms[MP3B_STAGE_H2D] += ms; }
      float ms = 0; cudaEventElapsedTime(&ms, b->evp[par][0], b->evp[par][7]); b->stage_ms[MP3B_STAGE_TOTAL] += ms;
    }
    const uint32_t *h_emit_n = b->h_emit_n + (size_t)par * S;
    const uint16_t *h_emit_size = b->h_emit_size + (size_t)par * S * (Fc + 1);
    size_t cmin = SIZE_MAX, cmax = 0, useful = 0;
    for (int s = 0; s < S; ++s) {
      uint32_t ne = h_emit_n[s];
      const uint16_t *sz = h_emit_size + (size_t)s * (Fc + 1);
      if (ne) b->frame_sizes[s].insert(b->frame_sizes[s].end(), sz, sz + ne);
      if (progressive && ne) {
        uint32_t add = 0;
        for (uint32_t k = 0; k < ne; ++k) add += sz[k];
        cmin = std::min<size_t>(cmin, opos[s]); opos[s] += add; cmax = std::max<size_t>(cmax, opos[s]); useful += add;
      }
    }
    if (progressive && useful) {
      cmin &= ~(size_t)15; cmax = std::min(round_up<size_t>(cmax, 16), b->pb.out_stride);
      if ((cmax - cmin) * (size_t)S > 3 * useful + (1u << 20)) progressive = false;   // ragged batch: one compact copy at the end instead
      else CU(cudaMemcpy2DAsync(b->h_out + cmin, b->pb.out_stride, b->pb.out + cmin, b->pb.out_stride, cmax - cmin, S, cudaMemcpyDeviceToHost, b->st_d2h));
    }
    if (b->trace) {
      const PassBuffers &pb = b->pb;
      const int ngc = 2 * cfg.channels;
      for (int s = 0; s < S; ++s) {
        const uint32_t nf = hplan[s].n_frames;
        for (uint32_t f = 0; f < nf; ++f) {
          const FrameRec &r = b->h_rec[(size_t)s * (Fc + 1) + 1 + f];
          mp3b_frame_record fr{};
          fr.bitrate_index = r.br_index; fr.padding = r.padding; fr.frame_size = r.slot + cfg.header_bytes; fr.main_data_size = r.slot;
          fr.main_data_begin = r.mdb; fr.reservoir_bits = r.reservoir_bits; fr.huff_bytes = r.huff_bytes; fr.ms = r.ms; fr.is_final = r.is_final;
          fr.frame_energy = r.frame_energy;
          b->tr_frames[s].push_back(fr);
          for (int j = 0; j < ngc; ++j) {
            const GcSide &g = r.gc[j];
            mp3b_gc_record q{};
            q.part23_length = g.part23; q.big_values = g.big_values; q.global_gain = g.global_gain; q.gain_used = g.gain_used + g.pad;
            q.block_type = g.block_type; q.subblock_gain[0] = g.sbg[0]; q.subblock_gain[1] = g.sbg[1]; q.subblock_gain[2] = g.sbg[2];
            q.region0 = g.region0; q.region1 = g.region1; q.preflag = g.preflag; q.g0 = g.g0; q.max_bits = g.max_bits;
            q.iterations = g.iterations; q.energy = g.energy;
            q.table_select[0] = g.tsel[0]; q.table_select[1] = g.tsel[1]; q.table_select[2] = g.tsel[2]; q.count1table_select = g.c1sel;
            q.scalefac_compress = g.sfc; q.part2_length = g.part2;
            b->tr_gc[s].push_back(q);
          }
        }
        const size_t cnt = (size_t)nf * ngc * 576, off = (size_t)s * b->GC * 576;
        if (cnt) {
          if (b->trace & 1) { size_t o = b->tr_spec[s].size(); b->tr_spec[s].resize(o + cnt); CU(cudaMemcpy(b->tr_spec[s].data() + o, pb.spec + off, cnt * 4, cudaMemcpyDeviceToHost)); }
          if (b->trace & 2) { size_t o = b->tr_ix[s].size(); b->tr_ix[s].resize(o + cnt); CU(cudaMemcpy(b->tr_ix[s].data() + o, pb.tr_ix + off, cnt * 4, cudaMemcpyDeviceToHost)); }
          if (b->trace & 4) { size_t o = b->tr_thr[s].size(); b->tr_thr[s].resize(o + cnt); CU(cudaMemcpy(b->tr_thr[s].data() + o, pb.tr_thr + off, cnt * 4, cudaMemcpyDeviceToHost)); }
          if (cfg.iso >= 2) {                                       // psychoacoustic ratios / PE and the scalefactors, 24 values per gc
            const size_t c24 = (size_t)nf * ngc * 24, o24 = (size_t)s * b->GC * 24;
            size_t o = b->tr_psy[s].size(); b->tr_psy[s].resize(o + c24);
            CU(cudaMemcpy(b->tr_psy[s].data() + o, pb.gc_psy + o24, c24 * 4, cudaMemcpyDeviceToHost));
            std::vector<uint8_t> tmp(c24);
            CU(cudaMemcpy(tmp.data(), pb.gc_sf + o24, c24, cudaMemcpyDeviceToHost));
            b->tr_sf[s].insert(b->tr_sf[s].end(), tmp.begin(), tmp.end());
          }
        }
      }
    }
    return MP3B_OK;
  };
  int slot = 0, par = 0;
  int open_slot = -1, open_par = 0;                              // the pass whose host-side work is still outstanding
  b->planned = true;                                             // from here on the host bookkeeping (pending) runs ahead of the device
  bool have = plan_pass(slot);
  if (have) { rc = issue_h2d(slot); if (rc) return rc; }
  while (have) {
    const StreamPlan *hplan = b->h_plan + (size_t)slot * S;
    CU(cudaStreamWaitEvent(st, b->ev_h2d[slot][1], 0));
    if (tail_used[par]) CU(cudaStreamWaitEvent(st, b->ev_tail_done[par], 0));   // the tail two passes back has let go of this parity's arrays
    cudaEvent_t *ev = b->evp[par];
    CU(cudaEventRecord(ev[0], st));
    CU(cudaEventRecord(ev[1], st));
    // ---- device pipeline
    PassBuffers pb = b->pb;
    if (par && overlap) {
      const PassBuffers &q = b->alt;
      pb.ms = q.ms; pb.frame_energy = q.frame_energy; pb.gc_energy = q.gc_energy; pb.gc_bt = q.gc_bt; pb.frame_br = q.frame_br; pb.smag = q.smag;
      pb.gc_meta = q.gc_meta; pb.gc_bits = q.gc_bits; pb.gc_bv = q.gc_bv;
      if (q.gc_psy) { pb.gc_psy = q.gc_psy; pb.gc_sf = q.gc_sf; }
    }
    pb.plan = b->d_plan[slot];
    pb.max_frames = 0;
    for (int s = 0; s < S; ++s) pb.max_frames = std::max<int>(pb.max_frames, (int)hplan[s].n_frames);
    pb.head_in = b->d_head[b->head_sel]; pb.head_out = b->d_head[b->head_sel ^ 1];
#define LAUNCH(expr)                                                                                        \
  do {                                                                                                      \
    int k = (expr);                                                                                             \
    if (k < 0) { b->sticky = MP3B_ERR_CUDA; return fail(MP3B_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString((cudaError_t)(-k))); } \
    b->launches += k;                                                                                       \
  } while (0)
    // CBR without joint stereo needs nothing from the pre-pass but the block type, which k_granule then derives from the PCM
    // itself (one pass less over the input); the trace plane keeps the pre-pass for its energy records
    const bool fused_prepass = !cfg.vbr && cfg.mode != 2 && !b->trace;
    // ---- head (stream st): everything that depends on the PCM alone
    if (!fused_prepass) LAUNCH(launch_prepass(cfg, pb, st));
    if (cfg.iso >= 3) LAUNCH(launch_blocktype(cfg, pb, st));       // ISO mode level 3: block types with one granule of look-ahead
    CU(cudaEventRecord(ev[2], st));
    LAUNCH(launch_spectrum(cfg, pb, st));
    CU(cudaEventRecord(ev[3], st));
    if (cfg.iso >= 2) LAUNCH(launch_psy(cfg, pb, st));             // ISO mode level 2: thresholds and PE from the PCM ...
    LAUNCH(launch_curve(cfg, pb, st, fused_prepass));
    if (cfg.iso >= 2) LAUNCH(launch_outer(cfg, pb, st));           // ... and the scalefactor outer loop on k_granule's magnitudes
    if (b->trace & 4) LAUNCH(launch_thresholds(cfg, pb, st));
    LAUNCH(launch_carry_head(cfg, pb, st));
    CU(cudaEventRecord(ev[9], st));
    CU(cudaEventRecord(b->ev_consumed[slot], st));                 // the staging buffer is free (the tail does not read PCM)
    CU(cudaEventRecord(b->ev_head_done[par], st));
    b->head_sel ^= 1;
    // ---- tail (stream stt): the serial scan and what hangs on it; in order on its stream, so scan(p + 1) follows carry_tail(p)
    CU(cudaStreamWaitEvent(stt, b->ev_head_done[par], 0));
    CU(cudaEventRecord(ev[4], stt));
    LAUNCH(launch_scan(cfg, pb, stt));
    if (cfg.iso) LAUNCH(launch_clear_md(cfg, pb, stt));            // stuffing bytes of the reservoir are never written: start from zeros
    CU(cudaEventRecord(ev[5], stt));
    LAUNCH(launch_pack(cfg, pb, stt));
    CU(cudaEventRecord(ev[6], stt));
    LAUNCH(launch_frames(cfg, pb, stt));
    LAUNCH(launch_carry_tail(cfg, pb, stt));
    CU(cudaEventRecord(ev[7], stt));
    CU(cudaEventRecord(b->ev_tail_done[par], stt));
    tail_used[par] = true;
    b->passes += 1;
    CU(cudaMemcpyAsync(b->h_emit_n + (size_t)par * S, pb.emit_n, (size_t)S * sizeof(uint32_t), cudaMemcpyDeviceToHost, stt));
    CU(cudaMemcpyAsync(b->h_emit_size + (size_t)par * S * (Fc + 1), pb.emit_size, (size_t)S * (Fc + 1) * sizeof(uint16_t), cudaMemcpyDeviceToHost, stt));
    if (b->trace) {
      b->h_rec.resize((size_t)S * (Fc + 1));
      CU(cudaMemcpyAsync(b->h_rec.data(), pb.rec, b->h_rec.size() * sizeof(FrameRec), cudaMemcpyDeviceToHost, stt));
    }
    CU(cudaEventRecord(ev[8], stt));
    // the previous pass: its kernels finished before this one's started, so this wait is short and the device stays busy
    if (open_slot >= 0) { rc = finish_pass(open_slot, open_par); open_slot = -1; if (rc) return rc; }
    open_slot = slot; open_par = par;
    if (b->trace) { rc = finish_pass(open_slot, open_par); open_slot = -1; if (rc) return rc; }
    // plan the next pass and start its transfer (its plan / staging slot belonged to the pass that was just finished)
    bool more = false;
    for (int s = 0; s < S && !more; ++s) more = cursor[s] < n[s] || (want_flush[s] && !flushed[s]);
    bool have_next = false;
    if (more) { have_next = plan_pass(slot ^ 1); if (have_next) { rc = issue_h2d(slot ^ 1); if (rc) return rc; } }
    have = have_next;
    slot ^= 1; par ^= 1;
  }
  if (open_slot >= 0) { rc = finish_pass(open_slot, open_par); if (rc) return rc; }
  // ---- results: counters, lengths, optional download
  if (overlap) for (int k = 0; k < 2; ++k) if (tail_used[k]) CU(cudaStreamWaitEvent(st, b->ev_tail_done[k], 0));
  CU(cudaEventRecord(b->ev[0], st));
  LAUNCH(launch_compact(cfg, b->pb, b->d_offsets, nullptr, 0, st));
  CU(cudaMemcpyAsync(b->h_state, b->pb.state, (size_t)S * sizeof(StreamState), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(b->h_offsets, b->d_offsets, (size_t)(S + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int err = 0;
  for (int s = 0; s < S; ++s) {
    const StreamState &hs = b->h_state[s];
    b->out_len[s] = hs.out_pos; b->frame_count[s] = hs.frame_count; b->byte_count[s] = hs.total_bytes;
    err |= hs.error;
  }
  b->out_total = 0;
  for (int s = 0; s < S; ++s) b->out_total += b->out_len[s];
  if (err) { b->sticky = MP3B_ERR_INTERNAL; return fail(MP3B_ERR_INTERNAL, "engine limit exceeded (flags 0x%x: 1 curve, 2 main-data buffer, 4 output buffer, 8 backlog > %d bytes, 16 ISO-mode bit count mismatch)", err, kMdCarryCap); }
  if (progressive) {
    CU(cudaStreamSynchronize(b->st_d2h));
    b->h_pitch = b->pb.out_stride;
    b->have_host_out = true;
  } else if (download) {
    CU(cudaStreamSynchronize(b->st_d2h));                 // a progressive attempt may still be copying into h_out
    const size_t total = (size_t)b->h_offsets[S];
    if (total > b->compact_cap) {
      if (b->d_compact) cudaFree(b->d_compact);
      b->d_compact = nullptr; b->compact_cap = 0;
      size_t cap = round_up<size_t>(total + total / 8 + 4096, 4096);
      CU(cudaMalloc((void **)&b->d_compact, cap));
      b->compact_cap = cap;
    }
    if (total > b->h_out_cap) {
      if (b->h_out) cudaFreeHost(b->h_out);
      b->h_out = nullptr; b->h_out_cap = 0;
      size_t cap = round_up<size_t>(total + total / 8 + 4096, 4096);
      CU(cudaHostAlloc((void **)&b->h_out, cap, cudaHostAllocDefault));
      b->h_out_cap = cap;
    }
    if (total) {
      LAUNCH(launch_compact(cfg, b->pb, b->d_offsets, b->d_compact, 1, st));
      CU(cudaMemcpyAsync(b->h_out, b->d_compact, total, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
    }
    b->have_host_out = true;
  }
  CU(cudaEventRecord(b->ev[1], st));
  CU(cudaEventSynchronize(b->ev[1]));
  { float ms = 0; cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]); b->stage_ms[MP3B_STAGE_D2H] += ms; b->stage_ms[MP3B_STAGE_TOTAL] += ms; }
  return MP3B_OK;
#undef LAUNCH
}

// A failure after the first plan_pass leaves the host-side bookkeeping (pending floats, cursors) ahead of the device state:
// the batch is then unusable until mp3b_batch_reset, whatever the failing call was.
int run_call(mp3b_batch *b, const float *const *pcm, const size_t *n_floats, bool device_ptrs, int flush,
             const uint8_t *flush_mask, bool download, size_t row_floats = 0, int elem_bytes = 4) {
  if (b) b->planned = false;
  const int rc = run_call_impl(b, pcm, n_floats, device_ptrs, flush, flush_mask, download, row_floats, elem_bytes);
  if (rc != MP3B_OK && b && b->planned && !b->sticky) b->sticky = rc;
  return rc;
}

int copy_out(const uint8_t *src, size_t len, uint8_t *out, size_t cap, size_t *written) {
  if (written) *written = len;
  if (len > cap) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "output needs %zu bytes, capacity is %zu", len, cap);
  if (len && !out) return fail(MP3B_ERR_BAD_ARG, "null output buffer");
  if (len) memcpy(out, src, len);
  return MP3B_OK;
}

// generateXingHeader SRC:367-420 + generateTOC SRC:423-449
int xing_header(const mp3b_batch *b, int stream, uint8_t *out, size_t cap, size_t *written) {
  const Config &c = b->cfg;
  const int frame_size = c.frame_base[c.cbr_index];
  std::vector<uint8_t> v;
  uint32_t h = 0;                                                 // SRC:379-392: no CRC, no padding, original = 1, copyright = 0
  h = 0x7FFu; h = h << 2 | 3u; h = h << 2 | 1u; h = h << 1 | 1u; h = h << 4 | (uint32_t)c.cbr_index; h = h << 2 | (uint32_t)c.sr_index;
  h = h << 1 | 0u; h = h << 1 | 0u; h = h << 2 | (uint32_t)c.mode_bits; h = h << 2 | (uint32_t)c.mode_ext; h = h << 1 | 0u; h = h << 1 | 1u; h = h << 2 | 0u;
  for (int i = 3; i >= 0; --i) v.push_back((uint8_t)(h >> (8 * i)));
  v.insert(v.end(), (size_t)c.side_bytes, 0);
  const char *tag = c.vbr ? "Xing" : "Info";
  v.insert(v.end(), tag, tag + 4);
  uint32_t words[3] = {0x07u, b->frame_count[stream] + 1u, b->byte_count[stream] + (uint32_t)frame_size};
  for (uint32_t w : words) for (int i = 3; i >= 0; --i) v.push_back((uint8_t)(w >> (8 * i)));
  const std::vector<uint16_t> &fs = b->frame_sizes[stream];
  long long total = 0;
  for (uint16_t z : fs) total += z;
  if (fs.empty() || total <= 0) { for (int p = 0; p < 100; ++p) v.push_back((uint8_t)(p * 255 / 99)); }
  else {
    std::vector<long long> cum(fs.size()); long long run = 0;
    for (size_t i = 0; i < fs.size(); ++i) { run += fs[i]; cum[i] = run; }
    for (int p = 0; p < 100; ++p) {
      size_t target = (size_t)p * fs.size() / 100;
      long long pos = target > 0 ? cum[target - 1] : 0, scaled = pos * 255 / total;
      v.push_back((uint8_t)std::min<long long>(scaled, 255));
    }
  }
  if ((int)v.size() < frame_size) v.resize((size_t)frame_size, 0);
  return copy_out(v.data(), v.size(), out, cap, written);
}

}  // namespace

extern "C" {

int mp3b_version(void) { return MP3B_VERSION; }
const char *mp3b_last_error(void) { return g_err.c_str(); }

void mp3b_options_default(mp3b_options *o) {                       // SRC:95-115
  if (!o) return;
  o->sample_rate = 44100; o->bitrate_kbps = 128; o->vbr = 0; o->mode = 1; o->quality = 5; o->crc_protected = 0; o->original = 1; o->copyright = 0;
}

int mp3b_device_count(int *count) {
  if (!count) return fail(MP3B_ERR_BAD_ARG, "null count");
  int n = 0; *count = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(MP3B_ERR_CUDA, "cudaGetDeviceCount failed"); }
  for (int i = 0; i < n; ++i) *count += usable_device(i);
  return MP3B_OK;
}

int mp3b_batch_create(const mp3b_options *opts, int n_streams, int device, mp3b_batch **out) { return create_batch(opts, n_streams, device, 0, out); }
int mp3b_batch_create_ex(const mp3b_options *opts, int n_streams, int device, int frames_per_pass, mp3b_batch **out) {
  if (frames_per_pass < 0 || frames_per_pass > 16384) return fail(MP3B_ERR_BAD_ARG, "frames_per_pass out of range");
  return create_batch(opts, n_streams, device, frames_per_pass, out);
}
// Multi-device batch: streams are cut into contiguous blocks, block k lives on devices[k] as a device batch of its own with a
// host thread of its own; no data ever moves between devices (SRC:237-258: sessions share nothing).  A device may be listed
// more than once (two independent pipelines on one GPU, or more streams than one device batch holds).
int mp3b_batch_create_multi(const mp3b_options *opts, int n_streams, const int *devices, int n_dev, int frames_per_pass, mp3b_batch **out) {
  if (!opts || !out || !devices || n_dev <= 0 || n_streams < n_dev) return fail(MP3B_ERR_BAD_ARG, "null pointer, n_dev <= 0 or fewer streams than devices");
  if (frames_per_pass < 0 || frames_per_pass > 16384) return fail(MP3B_ERR_BAD_ARG, "frames_per_pass out of range");
  Config cfg;
  int rc = fill_config(*opts, n_streams, cfg);
  if (rc) return rc;
  mp3b_batch *b = new mp3b_batch();
  b->S = n_streams; b->opt = *opts; b->cfg = cfg; b->ch = cfg.channels; b->device = devices[0];
  b->parts.assign((size_t)n_dev, nullptr);
  for (int k = 0; k <= n_dev; ++k) b->part_lo.push_back((int)((long long)n_streams * k / n_dev));
  for (int k = 0; k < n_dev; ++k) {
    b->workers.emplace_back(new PartWorker());
    b->workers.back()->th = std::thread(part_worker_main, b->workers.back().get());
  }
  const mp3b_options o = *opts;
  rc = on_all_parts(b->workers, [&](int k) { return create_batch(&o, b->part_lo[k + 1] - b->part_lo[k], devices[k], frames_per_pass, &b->parts[(size_t)k]); });
  if (rc) {
    std::string keep = g_err;
    b->parts.erase(std::remove(b->parts.begin(), b->parts.end(), nullptr), b->parts.end());
    if (b->parts.empty()) { stop_workers(b->workers); delete b; } else free_batch(b);
    g_err = keep;
    return rc;
  }
  b->Fc = b->parts[0]->Fc; b->GC = b->parts[0]->GC; b->max_frame_bytes = b->parts[0]->max_frame_bytes;
  *out = b;
  return MP3B_OK;
}
int mp3b_batch_device_count(const mp3b_batch *b) { return !b ? 0 : is_multi(b) ? (int)b->parts.size() : 1; }
void mp3b_batch_destroy(mp3b_batch *b) { free_batch(b); }
int mp3b_batch_stream_count(const mp3b_batch *b) { return b ? b->S : 0; }
int mp3b_batch_frames_per_pass(const mp3b_batch *b) { return b ? b->Fc : 0; }

// a call on a parent = the same call on every part, each on its own host thread; afterwards the parent carries the merged
// measurement fields (stage times: the slowest device; launches: all of them)
static int multi_call(mp3b_batch *b, const std::function<int(mp3b_batch *, int)> &fn) {
  if (b->sticky) return fail(b->sticky, "batch is in a failed state");
  const int rc = on_all_parts(b->workers, [&](int k) { return fn(b->parts[(size_t)k], b->part_lo[(size_t)k]); });
  if (rc) { std::string keep = g_err; b->sticky = rc; g_err = keep; return rc; }
  b->out_total = 0; b->launches = 0; b->passes = 0;
  for (auto &m : b->stage_ms) m = 0.0f;
  for (mp3b_batch *p : b->parts) {
    b->out_total += p->out_total; b->launches += p->launches; b->passes = std::max(b->passes, p->passes);
    for (int i = 0; i < MP3B_STAGE_COUNT; ++i) b->stage_ms[i] = std::max(b->stage_ms[i], p->stage_ms[i]);
  }
  return MP3B_OK;
}

int mp3b_batch_encode(mp3b_batch *b, const float *const *pcm, const size_t *n_floats, int flush, const uint8_t *flush_mask) {
  if (is_multi(b)) return multi_call(b, [&](mp3b_batch *p, int lo) { return run_call(p, pcm ? pcm + lo : nullptr, n_floats ? n_floats + lo : nullptr, false, flush, flush_mask ? flush_mask + lo : nullptr, true); });
  return run_call(b, pcm, n_floats, false, flush, flush_mask, true);
}
int mp3b_batch_encode_strided(mp3b_batch *b, const float *base, size_t pitch_floats, const size_t *n_floats, int flush, const uint8_t *flush_mask) {
  if (!b || !base || !n_floats) return fail(MP3B_ERR_BAD_ARG, "null batch / base / n_floats");
  if (is_multi(b)) return multi_call(b, [&](mp3b_batch *p, int lo) { return mp3b_batch_encode_strided(p, base + (size_t)lo * pitch_floats, pitch_floats, n_floats + lo, flush, flush_mask ? flush_mask + lo : nullptr); });
  std::vector<const float *> rows((size_t)b->S);
  for (int s = 0; s < b->S; ++s) {
    if (n_floats[s] > pitch_floats) return fail(MP3B_ERR_BAD_ARG, "stream %d: n_floats exceeds the row pitch", s);
    rows[(size_t)s] = base + (size_t)s * pitch_floats;
  }
  return run_call(b, rows.data(), n_floats, false, flush, flush_mask, true, pitch_floats);
}
int mp3b_batch_encode_i16(mp3b_batch *b, const int16_t *const *pcm, const size_t *n_samples, int flush, const uint8_t *flush_mask) {
  if (is_multi(b)) return multi_call(b, [&](mp3b_batch *p, int lo) { return mp3b_batch_encode_i16(p, pcm ? pcm + lo : nullptr, n_samples ? n_samples + lo : nullptr, flush, flush_mask ? flush_mask + lo : nullptr); });
  return run_call(b, reinterpret_cast<const float *const *>(pcm), n_samples, false, flush, flush_mask, true, 0, 2);
}
int mp3b_batch_encode_device(mp3b_batch *b, const float *const *d_pcm, const size_t *n_floats, int flush, int download) {
  // multi-device: d_pcm[i] must live on the device that owns stream i (mp3b_batch_stream_device)
  if (is_multi(b)) return multi_call(b, [&](mp3b_batch *p, int lo) { return run_call(p, d_pcm ? d_pcm + lo : nullptr, n_floats ? n_floats + lo : nullptr, true, flush, nullptr, download != 0); });
  return run_call(b, d_pcm, n_floats, true, flush, nullptr, download != 0);
}
int mp3b_batch_output(const mp3b_batch *b, int stream, const uint8_t **data, size_t *len) {
  if (!b || stream < 0 || stream >= b->S || !data || !len) return fail(MP3B_ERR_BAD_ARG, "bad stream index or null pointer");
  if (is_multi(b)) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_output(p, l, data, len); }
  if (!b->have_host_out) return fail(MP3B_ERR_BAD_ARG, "the last call did not download its output");
  *data = b->h_out ? b->h_out + (b->h_pitch ? (size_t)stream * b->h_pitch : (size_t)b->h_offsets[stream]) : nullptr; *len = b->out_len[stream];
  return MP3B_OK;
}
int mp3b_batch_output_device(const mp3b_batch *b, int stream, const uint8_t **d_data, size_t *len) {
  if (!b || stream < 0 || stream >= b->S || !d_data || !len) return fail(MP3B_ERR_BAD_ARG, "bad stream index or null pointer");
  if (is_multi(b)) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_output_device(p, l, d_data, len); }
  *d_data = b->pb.out + (size_t)stream * b->pb.out_stride; *len = b->out_len[stream];
  return MP3B_OK;
}
size_t mp3b_batch_output_total(const mp3b_batch *b) { return b ? b->out_total : 0; }
int mp3b_batch_xing_header(const mp3b_batch *b, int stream, uint8_t *out, size_t cap, size_t *written) {
  if (!b || stream < 0 || stream >= b->S) return fail(MP3B_ERR_BAD_ARG, "bad stream index");
  if (is_multi(b)) { int l; const mp3b_batch *p = part_of(b, stream, &l); return xing_header(p, l, out, cap, written); }
  return xing_header(b, stream, out, cap, written);
}
uint32_t mp3b_batch_frame_count(const mp3b_batch *b, int stream) {
  if (!b || stream < 0 || stream >= b->S) return 0;
  if (is_multi(b)) { int l; const mp3b_batch *p = part_of(b, stream, &l); return p->frame_count[l]; }
  return b->frame_count[stream];
}
uint32_t mp3b_batch_byte_count(const mp3b_batch *b, int stream) {
  if (!b || stream < 0 || stream >= b->S) return 0;
  if (is_multi(b)) { int l; const mp3b_batch *p = part_of(b, stream, &l); return p->byte_count[l]; }
  return b->byte_count[stream];
}
int mp3b_batch_stream_device(const mp3b_batch *b, int stream) {
  if (!b || stream < 0 || stream >= b->S) return fail(MP3B_ERR_BAD_ARG, "bad stream index");
  if (is_multi(b)) { int l; return part_of(b, stream, &l)->device; }
  return b->device;
}

// ---- session plane: a batch of one ------------------------------------------------------------------------
int mp3b_session_create(const mp3b_options *opts, int device, mp3b_session **out) {
  if (!out) return fail(MP3B_ERR_BAD_ARG, "null out");
  mp3b_batch *b = nullptr;
  int rc = create_batch(opts, 1, device, 1024, &b);                  // 1024 frames (27 s at 44.1 kHz) per pass: ~45 MB of device memory per session
  if (rc) return rc;
  mp3b_session *s = new mp3b_session(); s->b = b; *out = s;
  return MP3B_OK;
}
void mp3b_session_destroy(mp3b_session *s) { if (!s) return; free_batch(s->b); delete s; }

static int session_call(mp3b_session *s, const float *pcm, size_t n_floats, int flush, uint8_t *out, size_t cap, size_t *written) {
  if (!s) return fail(MP3B_ERR_BAD_ARG, "null session");
  if (!s->pending_out.empty()) return fail(MP3B_ERR_BAD_ARG, "output of the previous call is still pending: call mp3b_session_take_output");
  const float *ptrs[1] = {pcm}; size_t ns[1] = {pcm ? n_floats : 0};
  int rc = run_call(s->b, ptrs, ns, false, flush, nullptr, true);
  if (rc) return rc;
  const uint8_t *data; size_t len;
  rc = mp3b_batch_output(s->b, 0, &data, &len);
  if (rc) return rc;
  rc = copy_out(data, len, out, cap, written);
  if (rc == MP3B_ERR_BUFFER_TOO_SMALL) s->pending_out.assign(data, data + len);
  return rc;
}
int mp3b_session_encode(mp3b_session *s, const float *pcm, size_t n_floats, uint8_t *out, size_t cap, size_t *written) {
  return session_call(s, pcm, n_floats, 0, out, cap, written);
}
int mp3b_session_flush(mp3b_session *s, uint8_t *out, size_t cap, size_t *written) { return session_call(s, nullptr, 0, 1, out, cap, written); }
int mp3b_session_take_output(mp3b_session *s, uint8_t *out, size_t cap, size_t *written) {
  if (!s) return fail(MP3B_ERR_BAD_ARG, "null session");
  int rc = copy_out(s->pending_out.data(), s->pending_out.size(), out, cap, written);
  if (rc == MP3B_OK) s->pending_out.clear();
  return rc;
}
size_t mp3b_session_output_bound(const mp3b_session *s, size_t n_floats) {
  if (!s) return 0;
  const mp3b_batch *b = s->b;
  return ((b->pending[0] + n_floats) / b->cfg.fsc + 2 + (b->cfg.iso >= 3 ? 1 : 0)) * (size_t)b->max_frame_bytes;
}
int mp3b_session_xing_header(const mp3b_session *s, uint8_t *out, size_t cap, size_t *written) {
  if (!s) return fail(MP3B_ERR_BAD_ARG, "null session");
  return xing_header(s->b, 0, out, cap, written);
}
int mp3b_xing_frame_size(const mp3b_options *o) {                  // SRC:198-200
  if (!o || o->sample_rate <= 0) return fail(MP3B_ERR_BAD_ARG, "null options or sample_rate <= 0");
  return (int)(144LL * bitrate_value(bitrate_index(o->bitrate_kbps, o->sample_rate)) * 1000 / o->sample_rate);
}
uint32_t mp3b_session_frame_count(const mp3b_session *s) { return s ? s->b->frame_count[0] : 0; }
uint32_t mp3b_session_byte_count(const mp3b_session *s) { return s ? s->b->byte_count[0] : 0; }

// ID3TagWriter.build SRC:1040-1075 (host only, no device involved)
int mp3b_id3_build(const mp3b_id3 *tag, uint8_t *out, size_t cap, size_t *written) {
  if (!tag) return fail(MP3B_ERR_BAD_ARG, "null tag");
  std::vector<uint8_t> f;
  auto frame_header = [&](const char *id, uint32_t size) {        // SRC:1127-1135
    f.insert(f.end(), id, id + 4);
    for (int i = 3; i >= 0; --i) f.push_back((uint8_t)(size >> (8 * i)));
    f.push_back(0); f.push_back(0);
  };
  auto text = [&](const char *id, const std::string &v) {         // SRC:1078-1086
    frame_header(id, (uint32_t)(1 + v.size())); f.push_back(0x03); f.insert(f.end(), v.begin(), v.end());
  };
  if (tag->title) text("TIT2", tag->title);
  if (tag->artist) text("TPE1", tag->artist);
  if (tag->album) text("TALB", tag->album);
  if (tag->genre) text("TCON", tag->genre);
  if (tag->year >= 0) text("TYER", std::to_string(tag->year));
  if (tag->track >= 0) text("TRCK", tag->track_total >= 0 ? std::to_string(tag->track) + "/" + std::to_string(tag->track_total) : std::to_string(tag->track));
  if (tag->comment) {                                              // SRC:1089-1099
    std::string c = tag->comment;
    frame_header("COMM", (uint32_t)(1 + 3 + 1 + c.size()));
    f.push_back(0x03); f.push_back('e'); f.push_back('n'); f.push_back('g'); f.push_back(0); f.insert(f.end(), c.begin(), c.end());
  }
  if (tag->album_art) {                                            // SRC:1102-1114
    std::string mime = tag->album_art_mime ? tag->album_art_mime : "image/jpeg";
    frame_header("APIC", (uint32_t)(1 + mime.size() + 1 + 1 + 1 + tag->album_art_len));
    f.push_back(0x03); f.insert(f.end(), mime.begin(), mime.end()); f.push_back(0); f.push_back(0x03); f.push_back(0);
    f.insert(f.end(), tag->album_art, tag->album_art + tag->album_art_len);
  }
  if (f.empty()) { if (written) *written = 0; return MP3B_OK; }   // SRC:1066
  std::vector<uint8_t> o = {0x49, 0x44, 0x33, 0x03, 0x00, 0x00};
  uint32_t sz = (uint32_t)f.size();
  o.push_back((sz >> 21) & 0x7F); o.push_back((sz >> 14) & 0x7F); o.push_back((sz >> 7) & 0x7F); o.push_back(sz & 0x7F);
  o.insert(o.end(), f.begin(), f.end());
  return copy_out(o.data(), o.size(), out, cap, written);
}

// ---- memory helpers ---------------------------------------------------------------------------------------
int mp3b_host_alloc(size_t bytes, void **out) {
  if (!out) return fail(MP3B_ERR_BAD_ARG, "null out");
  // MP3B_HOST_WC=1: write-combined pinned memory for upload buffers the host only writes (not snooped during DMA)
  static const bool wc = [] { const char *v = getenv("MP3B_HOST_WC"); return v && v[0] == '1'; }();
  CU(cudaHostAlloc(out, std::max<size_t>(bytes, 1), wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return MP3B_OK;
}
void mp3b_host_free(void *p) { if (p) cudaFreeHost(p); }
int mp3b_device_alloc(int device, size_t bytes, void **out) {
  if (!out) return fail(MP3B_ERR_BAD_ARG, "null out");
  CU(cudaSetDevice(device));
  CU(cudaMalloc(out, std::max<size_t>(bytes, 1)));
  return MP3B_OK;
}
void mp3b_device_free(int device, void *p) { if (p) { cudaSetDevice(device); cudaFree(p); } }
int mp3b_device_copy(int device, void *dst, const void *src, size_t bytes, int kind) {
  CU(cudaSetDevice(device));
  CU(cudaMemcpy(dst, src, bytes, kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
  return MP3B_OK;
}
int mp3b_device_sync(int device) { CU(cudaSetDevice(device)); CU(cudaDeviceSynchronize()); return MP3B_OK; }

// ---- measurement / test plane -------------------------------------------------------------------------------
int mp3b_batch_stage_ms(const mp3b_batch *b, float *ms, int n) {
  if (!b || !ms) return fail(MP3B_ERR_BAD_ARG, "null pointer");
  for (int i = 0; i < n && i < MP3B_STAGE_COUNT; ++i) ms[i] = b->stage_ms[i];
  return MP3B_STAGE_COUNT;
}
int mp3b_batch_launch_count(const mp3b_batch *b) { return b ? b->launches : 0; }
int mp3b_batch_pass_count(const mp3b_batch *b) { return b ? b->passes : 0; }
void *mp3b_batch_stream(const mp3b_batch *b) { return !b ? nullptr : is_multi(b) ? (void *)b->parts[0]->st : (void *)b->st; }
int mp3b_batch_reset(mp3b_batch *b) {
  if (!b) return fail(MP3B_ERR_BAD_ARG, "null batch");
  if (is_multi(b)) {
    const int rc = on_all_parts(b->workers, [&](int k) { return mp3b_batch_reset(b->parts[(size_t)k]); });
    if (rc == MP3B_OK) { b->sticky = 0; b->out_total = 0; }
    return rc;
  }
  CU(cudaSetDevice(b->device));
  const size_t S = b->S;
  CU(cudaMemsetAsync(b->pb.state, 0, S * sizeof(StreamState), b->st));
  CU(cudaMemsetAsync(b->d_head[0], 0, S * 2 * b->cfg.fsc * sizeof(float), b->st));
  CU(cudaMemsetAsync(b->d_head[1], 0, S * 2 * b->cfg.fsc * sizeof(float), b->st));
  CU(cudaMemset2DAsync(b->pb.sub, (size_t)b->pb.sub_rows * 32 * sizeof(float), 0, 576 * sizeof(float), S * b->cfg.channels, b->st));
  CU(cudaStreamSynchronize(b->st));
  std::fill(b->pending.begin(), b->pending.end(), 0u); std::fill(b->fed.begin(), b->fed.end(), (uint8_t)0);
  std::fill(b->out_len.begin(), b->out_len.end(), 0u);
  std::fill(b->frame_count.begin(), b->frame_count.end(), 0u);
  std::fill(b->byte_count.begin(), b->byte_count.end(), 0u);
  for (auto &v : b->frame_sizes) v.clear();
  b->out_total = 0; b->have_host_out = false; b->sticky = 0;
  return MP3B_OK;
}
// EncoderSession is a value type in the reference (SRC:237-258: every field is a value): copying the struct is a full
// snapshot of the encoder.  The same here: a new batch with the same options and a copy of everything that crosses a call
// boundary — per-stream device state (reservoir, padding remainder, counters, VBR history, buffered frame), the carried PCM
// frame, the MDCT overlap rows, the main-data backlog — and of the host-side counters.
int mp3b_batch_clone(const mp3b_batch *src, mp3b_batch **out) {
  if (!src || !out) return fail(MP3B_ERR_BAD_ARG, "null batch / out");
  if (src->sticky) return fail(src->sticky, "cannot clone a batch in a failed state");
  if (is_multi(src)) {                                            // part by part, each on a host thread of the clone
    mp3b_batch *c = new mp3b_batch();
    c->S = src->S; c->opt = src->opt; c->cfg = src->cfg; c->ch = src->ch; c->device = src->device; c->Fc = src->Fc; c->GC = src->GC;
    c->max_frame_bytes = src->max_frame_bytes; c->part_lo = src->part_lo;
    c->parts.assign(src->parts.size(), nullptr);
    for (size_t k = 0; k < src->parts.size(); ++k) {
      c->workers.emplace_back(new PartWorker());
      c->workers.back()->th = std::thread(part_worker_main, c->workers.back().get());
    }
    const int rc = on_all_parts(c->workers, [&](int k) { return mp3b_batch_clone(src->parts[(size_t)k], &c->parts[(size_t)k]); });
    if (rc) {
      std::string keep = g_err;
      c->parts.erase(std::remove(c->parts.begin(), c->parts.end(), nullptr), c->parts.end());
      if (c->parts.empty()) { stop_workers(c->workers); delete c; } else free_batch(c);
      g_err = keep;
      return rc;
    }
    *out = c;
    return MP3B_OK;
  }
  mp3b_batch *b = nullptr;
  int rc = create_batch(&src->opt, src->S, src->device, src->Fc, &b);
  if (rc) return rc;
  const size_t S = (size_t)src->S, fsc2 = 2 * (size_t)src->cfg.fsc, ch = (size_t)src->cfg.channels;
  cudaError_t e = cudaStreamSynchronize(src->st);
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  // on the clone's own stream (create_batch has already waited for its zero fills), finished before the handle is returned
  A(cudaMemcpyAsync(b->pb.state, src->pb.state, S * sizeof(StreamState), cudaMemcpyDeviceToDevice, b->st));
  A(cudaMemcpyAsync(b->d_head[0], src->d_head[0], S * fsc2 * sizeof(float), cudaMemcpyDeviceToDevice, b->st));
  A(cudaMemcpyAsync(b->d_head[1], src->d_head[1], S * fsc2 * sizeof(float), cudaMemcpyDeviceToDevice, b->st));
  A(cudaMemcpy2DAsync(b->pb.sub, (size_t)b->pb.sub_rows * 32 * sizeof(float), src->pb.sub, (size_t)src->pb.sub_rows * 32 * sizeof(float),
                      576 * sizeof(float), S * ch, cudaMemcpyDeviceToDevice, b->st));
  A(cudaMemcpyAsync(b->pb.md_carry, src->pb.md_carry, S * kMdCarryCap, cudaMemcpyDeviceToDevice, b->st));
  A(cudaStreamSynchronize(b->st));
  if (e != cudaSuccess) { free_batch(b); return fail(MP3B_ERR_CUDA, "clone failed: %s", cudaGetErrorString(e)); }
  b->head_sel = src->head_sel; b->cfg.iso = src->cfg.iso; b->cfg.ms_scale = src->cfg.ms_scale; b->cfg.iso_delay = src->cfg.iso_delay;
  if (b->cfg.iso >= 2 && ensure_iso2(b) != MP3B_OK) { free_batch(b); return MP3B_ERR_CUDA; }
  if (src->pb.tc_b && mp3b_batch_set_matrixing(b, 1) != MP3B_OK) { free_batch(b); return MP3B_ERR_CUDA; }
  b->pending = src->pending; b->fed = src->fed; b->frame_count = src->frame_count; b->byte_count = src->byte_count; b->frame_sizes = src->frame_sizes;
  b->trace = src->trace;
  *out = b;
  return MP3B_OK;
}
int mp3b_session_clone(const mp3b_session *s, mp3b_session **out) {
  if (!s || !out) return fail(MP3B_ERR_BAD_ARG, "null session / out");
  mp3b_batch *b = nullptr;
  int rc = mp3b_batch_clone(s->b, &b);
  if (rc) return rc;
  mp3b_session *c = new mp3b_session(); c->b = b; c->pending_out = s->pending_out; *out = c;
  return MP3B_OK;
}
int mp3b_batch_reset_stream(mp3b_batch *b, int stream) {
  if (!b || stream < 0 || stream >= b->S) return fail(MP3B_ERR_BAD_ARG, "bad batch / stream");
  if (is_multi(b)) { int l; mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_reset_stream(p, l); }
  CU(cudaSetDevice(b->device));
  const size_t s = (size_t)stream, fsc2 = 2 * (size_t)b->cfg.fsc, ch = (size_t)b->cfg.channels;
  CU(cudaMemsetAsync(b->pb.state + s, 0, sizeof(StreamState), b->st));
  CU(cudaMemsetAsync(b->d_head[0] + s * fsc2, 0, fsc2 * sizeof(float), b->st));
  CU(cudaMemsetAsync(b->d_head[1] + s * fsc2, 0, fsc2 * sizeof(float), b->st));
  CU(cudaMemset2DAsync(b->pb.sub + s * ch * (size_t)b->pb.sub_rows * 32, (size_t)b->pb.sub_rows * 32 * sizeof(float), 0, 576 * sizeof(float), ch, b->st));
  CU(cudaStreamSynchronize(b->st));
  b->pending[s] = 0; b->fed[s] = 0; b->out_len[s] = 0; b->frame_count[s] = 0; b->byte_count[s] = 0; b->frame_sizes[s].clear();
  return MP3B_OK;
}
// Opt-in ISO mode (iso_mode.cuh).  Only on fresh sessions: the two modes do not share reservoir semantics.
int mp3b_batch_set_iso_mode(mp3b_batch *b, int on) {
  if (!b) return fail(MP3B_ERR_BAD_ARG, "null batch");
  if (is_multi(b)) {
    for (mp3b_batch *p : b->parts) { const int rc = mp3b_batch_set_iso_mode(p, on); if (rc) return rc; }
    b->cfg.iso = on < 0 ? 0 : on > 3 ? 3 : on;
    return MP3B_OK;
  }
  for (int s = 0; s < b->S; ++s)
    if (b->pending[(size_t)s] || b->frame_count[(size_t)s]) return fail(MP3B_ERR_BAD_ARG, "iso mode can only be changed on fresh sessions (after create or reset)");
  const int level = on < 0 ? 0 : on > 3 ? 3 : on;
  if (level >= 2) { const int rc = ensure_iso2(b); if (rc) return rc; }
  b->cfg.iso = level;
  b->cfg.iso_delay = level >= 3 ? 576 : 0;
  b->cfg.ms_scale = on ? 0.70710678118654752440f : 0.5f;
  return MP3B_OK;
}
int mp3b_batch_iso_mode(const mp3b_batch *b) { return b ? b->cfg.iso : 0; }
// Matrixing of the filterbank: 0 = FP32 FMA in the reference's order (default, bit-exact with the oracle), 1 = tensor cores with
// a three-term TF32 split (filterbank_tc.cuh).  May be switched at any time: it changes how S = M Y is summed, nothing else.
int mp3b_batch_set_matrixing(mp3b_batch *b, int mode) {
  if (!b) return fail(MP3B_ERR_BAD_ARG, "null batch");
  if (mode != 0 && mode != 1) return fail(MP3B_ERR_BAD_ARG, "matrixing must be 0 (FP32 FMA) or 1 (3xTF32 on the tensor cores)");
  if (is_multi(b)) {
    for (mp3b_batch *p : b->parts) { const int rc = mp3b_batch_set_matrixing(p, mode); if (rc) return rc; }
    return MP3B_OK;
  }
  if (mode == 1 && !b->d_tc_b) {
    CU(cudaSetDevice(b->device));
    // [2 n halves][96 rows = hi | mid | lo term of the 32 subbands][32 n], rows of 128 bytes in 1024-byte atoms whose 16-byte
    // chunks are XOR-swizzled with the row number: the shared-memory image of a K-major SWIZZLE_128B operand
    std::vector<float> img(2 * 96 * 32);
    const uint32_t mask = 0xFFFFE000u;
    for (int H = 0; H < 2; ++H)
      for (int r = 0; r < 96; ++r)
        for (int k = 0; k < 32; ++k) {
          const float m = tab::kAnalysis[r & 31][32 * H + k];
          uint32_t u; float hi, mid, lo, rest;
          memcpy(&u, &m, 4); u &= mask; memcpy(&hi, &u, 4);
          rest = m - hi; memcpy(&u, &rest, 4); u &= mask; memcpy(&mid, &u, 4);
          lo = rest - mid;
          const size_t off = (size_t)H * 96 * 128 + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 + (size_t)((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2));
          img[off / 4] = r < 32 ? hi : r < 64 ? mid : lo;
        }
    CU(cudaMalloc((void **)&b->d_tc_b, img.size() * sizeof(float)));
    CU(cudaMemcpy(b->d_tc_b, img.data(), img.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaStreamSynchronize(cudaStreamLegacy));
  }
  b->pb.tc_b = mode ? b->d_tc_b : nullptr;
  return MP3B_OK;
}
int mp3b_batch_matrixing(const mp3b_batch *b) { return b ? (is_multi(b) ? mp3b_batch_matrixing(b->parts[0]) : b->pb.tc_b != nullptr) : 0; }
int mp3b_session_set_iso_mode(mp3b_session *s, int on) { return s ? mp3b_batch_set_iso_mode(s->b, on) : fail(MP3B_ERR_BAD_ARG, "null session"); }
int mp3b_batch_set_trace(mp3b_batch *b, int flags) {
  if (!b) return fail(MP3B_ERR_BAD_ARG, "null batch");
  for (mp3b_batch *p : b->parts) p->trace = flags ? (flags | 8) : 0;
  b->trace = flags ? (flags | 8) : 0;
  return MP3B_OK;
}
int mp3b_batch_trace_frames(const mp3b_batch *b, int stream) {
  if (is_multi(b) && stream >= 0 && stream < b->S) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_trace_frames(p, l); }
  if (!b || stream < 0 || stream >= b->S || (size_t)stream >= b->tr_frames.size()) return 0;
  return (int)b->tr_frames[stream].size();
}
int mp3b_batch_trace_frame_records(const mp3b_batch *b, int stream, mp3b_frame_record *out, int cap) {
  if (is_multi(b) && stream >= 0 && stream < b->S) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_trace_frame_records(p, l, out, cap); }
  int n = mp3b_batch_trace_frames(b, stream);
  if (n > cap) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "need %d records", n);
  if (n) memcpy(out, b->tr_frames[stream].data(), (size_t)n * sizeof *out);
  return n;
}
int mp3b_batch_trace_gc_records(const mp3b_batch *b, int stream, mp3b_gc_record *out, int cap) {
  if (is_multi(b) && stream >= 0 && stream < b->S) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_trace_gc_records(p, l, out, cap); }
  if (!b || stream < 0 || (size_t)stream >= b->tr_gc.size()) return 0;
  int n = (int)b->tr_gc[stream].size();
  if (n > cap) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "need %d records", n);
  if (n) memcpy(out, b->tr_gc[stream].data(), (size_t)n * sizeof *out);
  return n;
}
int mp3b_batch_trace_gc_array(const mp3b_batch *b, int stream, int kind, void *out, int cap_gc) {
  if (is_multi(b) && stream >= 0 && stream < b->S) { int l; const mp3b_batch *p = part_of(b, stream, &l); return mp3b_batch_trace_gc_array(p, l, kind, out, cap_gc); }
  if (!b || stream < 0 || (size_t)stream >= b->tr_gc.size() || !out) return fail(MP3B_ERR_BAD_ARG, "bad stream or null out");
  const void *src; size_t elems;
  if (kind == 0) { src = b->tr_spec[stream].data(); elems = b->tr_spec[stream].size(); }
  else if (kind == 1) { src = b->tr_ix[stream].data(); elems = b->tr_ix[stream].size(); }
  else if (kind == 2) { src = b->tr_thr[stream].data(); elems = b->tr_thr[stream].size(); }
  else if (kind == 3 || kind == 4) {                               // 24 values per gc: psychoacoustic record f32 / scalefactor record i32
    const size_t n24 = kind == 3 ? b->tr_psy[stream].size() : b->tr_sf[stream].size();
    if (n24 > (size_t)cap_gc * 24) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "need %zu granule-channels", n24 / 24);
    if (n24) memcpy(out, kind == 3 ? (const void *)b->tr_psy[stream].data() : (const void *)b->tr_sf[stream].data(), n24 * 4);
    return (int)(n24 / 24);
  }
  else return fail(MP3B_ERR_BAD_ARG, "kind must be 0 ... 4");
  if (elems > (size_t)cap_gc * 576) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "need %zu granule-channels", elems / 576);
  if (elems) memcpy(out, src, elems * 4);
  return (int)(elems / 576);
}

int mp3b_table(int which, void *out, size_t cap_bytes) {
  const void *src = nullptr; size_t n = 0, es = 4;
  switch (which) {
    case 0: src = tab::kWindow; n = 512; break;
    case 1: src = tab::kAnalysis; n = 2048; break;
    case 2: src = tab::kMdctLong; n = 648; break;
    case 3: src = tab::kMdctShort; n = 72; break;
    case 4: src = tab::kWinLong; n = 36; break;
    case 5: src = tab::kWinShort; n = 12; break;
    case 6: src = host_inv_step(); n = 256; break;
    case 7: src = tab::kHuff15Len; n = 256; es = 1; break;
    case 8: src = tab::kHuff15Code; n = 256; es = 1; break;
    case 9: src = host_gain_thr(); n = 256; es = 8; break;
    case 10: src = tab::kAliasCs; n = 8; break;
    case 11: src = tab::kAliasCa; n = 8; break;
    case 12: src = host_sfb_cum(); n = 63; break;
    case 13: src = tab::kLen31s; n = 31 * 32; es = 1; break;
    case 14: src = tab::kTab31; n = 31 * 32; es = 2; break;
    default: return fail(MP3B_ERR_BAD_ARG, "unknown table %d", which);
  }
  if (n * es > cap_bytes) return fail(MP3B_ERR_BUFFER_TOO_SMALL, "table %d needs %zu bytes", which, n * es);
  if (out) memcpy(out, src, n * es);
  return (int)n;
}

int mp3b_synth_fill(int device, float *d_pcm, size_t n_samples_per_channel, int channels, int sample_rate, float f_left,
                    float f_right, float amp, float noise, uint64_t seed) {
  if (!d_pcm || channels < 1 || channels > 2 || sample_rate <= 0) return fail(MP3B_ERR_BAD_ARG, "bad synth arguments");
  CU(cudaSetDevice(device));
  int k = launch_synth(d_pcm, n_samples_per_channel, channels, sample_rate, f_left, f_right, amp, noise, seed, nullptr);
  if (k < 0) return fail(MP3B_ERR_CUDA, "synth launch failed: %s", cudaGetErrorString((cudaError_t)(-k)));
  CU(cudaDeviceSynchronize());
  return MP3B_OK;
}

int mp3b_selftest(int device, uint64_t mismatches[3]) {
  if (!mismatches) return fail(MP3B_ERR_BAD_ARG, "null mismatches");
  if (!usable_device(device)) return fail(MP3B_ERR_CUDA, "device %d is not compute capability 10.x", device);
  CU(cudaSetDevice(device));
  unsigned long long *d = nullptr;
  CU(cudaMalloc((void **)&d, 3 * sizeof *d));
  CU(cudaMemset(d, 0, 3 * sizeof *d));
  int k = launch_selftest(d, nullptr);
  if (k < 0) { cudaFree(d); return fail(MP3B_ERR_CUDA, "selftest launch failed: %s", cudaGetErrorString((cudaError_t)(-k))); }
  unsigned long long h[3] = {0, 0, 0};
  cudaError_t ce = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (ce != cudaSuccess) return fail(MP3B_ERR_CUDA, "selftest failed: %s", cudaGetErrorString(ce));
  for (int i = 0; i < 3; ++i) mismatches[i] = h[i];
  return MP3B_OK;
}

}  // extern "C"
