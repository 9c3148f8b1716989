// Session pool: N independent EncoderSessions (SRC:237-350) that are driven from many threads, one blocking
// encode(samples:) / flush() call per session at a time, and advanced on the GPU together.  This is the host-side
// coalescing layer of BASELINE config 5 ("1024 concurrent sessions fed 1152-sample chunks"): the calls that arrive within
// a short window become ONE step of the batch plane — one upload, one pass of the kernels, one download — instead of N
// launch-bound single-stream calls.  A step fires as soon as every session that is currently inside a call has deposited
// its request and either all open sessions are present or `max_wait_us` has elapsed since the first of them arrived.
//
// The worker thread owns the batch (and the CUDA work); client threads only copy their PCM in and their bytes out.
//
// Contract details: (1) a session that is open but not inside a call is indistinguishable from one that is about to call, so
// every step then waits the full `max_wait_us` for it — close (flush) sessions that go idle, or pick max_wait_us as the
// latency an idle peer may cost.  (2) A CUDA / engine-limit failure of a step is sticky for the batch (include/mp3b200.h):
// every session then fails until all of them have been closed; the next mp3b_pool_open resets the batch (fresh sessions).
// (3) mp3b_pool_destroy may be called while clients are blocked: they return MP3B_ERR_INTERNAL, and destroy waits for the
// last of them to leave the pool before it frees anything.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mp3b200.h"

namespace {
struct Slot {
  bool open = false;         // acquired by a client
  bool pending = false;      // request deposited, step not yet run
  bool done = false;         // result ready for the waiting client
  bool undelivered = false;  // the caller's buffer was too small: `out` waits for mp3b_pool_take_output
  int flush = 0, rc = 0;
  size_t n = 0;              // request: n floats, copied into the slot's arena row (the caller's array may be reused
                             // after return, SRC:298) or, when larger than a row, into `big`
  std::vector<float> big;
  std::vector<uint8_t> out;  // response
  std::string err;
  std::mutex m;              // guards `done`: the worker wakes exactly the members of a step, not every waiting client
  std::condition_variable cv;
};
}  // namespace

struct mp3b_pool {
  mp3b_batch *batch = nullptr;
  int n = 0, max_wait_us = 0;
  std::vector<std::unique_ptr<Slot>> slot_store;
  std::vector<Slot *> slots;
  std::mutex mu;
  std::condition_variable cv_worker;
  std::thread worker;
  bool stop = false;
  bool failed = false;       // a step left the batch in its sticky failed state
  int n_open = 0, n_pending = 0;
  int in_flight = 0;         // client threads currently inside pool_call / take_output
  std::condition_variable cv_idle;
  std::chrono::steady_clock::time_point first_pending;
  // statistics
  uint64_t steps = 0, requests = 0;
  // step scratch (worker only)
  float *arena = nullptr;    // pinned [n][row] staging: one strided upload per step
  size_t row = 0;
  std::vector<const float *> ptrs;
  std::vector<size_t> lens;
  std::vector<uint8_t> mask;
  std::vector<int> members;
};

static void pool_worker(mp3b_pool *p) {
  std::unique_lock<std::mutex> lk(p->mu);
  for (;;) {
    p->cv_worker.wait(lk, [&] { return p->stop || p->n_pending > 0; });
    if (p->stop) return;
    // coalescing window: until every open session has a request in, or the deadline of the first request passes
    const auto deadline = p->first_pending + std::chrono::microseconds(p->max_wait_us);
    while (!p->stop && p->n_pending < p->n_open) {
      if (p->cv_worker.wait_until(lk, deadline) == std::cv_status::timeout) break;
    }
    if (p->stop) return;
    // take the step's members; requests that arrive from now on belong to the next step
    p->members.clear();
    int any_flush = 0; bool in_arena = true;
    for (int i = 0; i < p->n; ++i) {
      Slot &s = *p->slots[i];
      const bool in = s.pending;
      p->ptrs[i] = !in ? nullptr : s.n > p->row ? s.big.data() : p->arena + (size_t)i * p->row;
      p->lens[i] = in ? s.n : 0;
      if (in && s.n > p->row) in_arena = false;
      p->mask[i] = in && s.flush ? 1 : 0;
      any_flush |= p->mask[i];
      if (in) { p->members.push_back(i); s.pending = false; }
    }
    p->n_pending = 0;
    lk.unlock();
    int rc = in_arena ? mp3b_batch_encode_strided(p->batch, p->arena, p->row, p->lens.data(), any_flush, any_flush ? p->mask.data() : nullptr)
                      : mp3b_batch_encode(p->batch, p->ptrs.data(), p->lens.data(), any_flush, any_flush ? p->mask.data() : nullptr);
    std::string err = rc ? mp3b_last_error() : "";
    for (int i : p->members) {
      Slot &s = *p->slots[i];                      // only this thread and the (blocked) owner touch a member slot now
      s.rc = rc; s.err = err; s.out.clear();
      if (rc == MP3B_OK) {
        const uint8_t *data = nullptr; size_t len = 0;
        if (mp3b_batch_output(p->batch, i, &data, &len) == MP3B_OK && len) s.out.assign(data, data + len);
        if (s.flush) mp3b_batch_reset_stream(p->batch, i);           // the slot goes back to a fresh EncoderSession
      }
    }
    for (int i : p->members) {
      Slot &s = *p->slots[i];
      { std::lock_guard<std::mutex> g(s.m); s.done = true; }
      s.cv.notify_one();
    }
    lk.lock();
    if (rc == MP3B_ERR_CUDA || rc == MP3B_ERR_INTERNAL || rc == MP3B_ERR_OOM) p->failed = true;
    p->steps += 1; p->requests += p->members.size();
  }
}

static thread_local std::string t_pool_err;

extern "C" {

int mp3b_pool_create(const mp3b_options *opts, int n_sessions, int device, int max_wait_us, mp3b_pool **out) {
  if (!out || n_sessions <= 0 || max_wait_us < 0) return MP3B_ERR_BAD_ARG;
  mp3b_batch *b = nullptr;
  int rc = mp3b_batch_create_ex(opts, n_sessions, device, 8, &b);      // short passes: the pool is for chunked streaming
  if (rc) return rc;
  mp3b_pool *p = new mp3b_pool();
  p->batch = b; p->n = n_sessions; p->max_wait_us = max_wait_us;
  for (int i = 0; i < n_sessions; ++i) { p->slot_store.emplace_back(new Slot()); p->slots.push_back(p->slot_store.back().get()); }
  p->ptrs.resize(n_sessions); p->lens.resize(n_sessions); p->mask.resize(n_sessions);
  p->row = 4 * 1152 * 2;                                               // four stereo frames per call fit a row
  void *arena = nullptr;
  rc = mp3b_host_alloc((size_t)n_sessions * p->row * sizeof(float), &arena);
  if (rc) { mp3b_batch_destroy(b); delete p; return rc; }
  p->arena = (float *)arena;
  p->worker = std::thread(pool_worker, p);
  *out = p;
  return MP3B_OK;
}

void mp3b_pool_destroy(mp3b_pool *p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; }
  p->cv_worker.notify_all();
  if (p->worker.joinable()) p->worker.join();       // a step in progress completes and delivers; nothing new starts
  for (Slot *s : p->slots) { { std::lock_guard<std::mutex> g(s->m); s->done = true; } s->cv.notify_all(); }
  {                                                  // blocked clients wake, see `stop`, and leave without touching the pool again
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_idle.wait(lk, [&] { return p->in_flight == 0; });
  }
  mp3b_batch_destroy(p->batch);
  mp3b_host_free(p->arena);
  delete p;
}

int mp3b_pool_open(mp3b_pool *p, int *slot) {
  if (!p || !slot) return MP3B_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(p->mu);
  if (p->stop) return MP3B_ERR_BAD_ARG;
  if (p->failed && p->n_open == 0) {                // every session of the failed batch is closed: start over with fresh ones
    if (mp3b_batch_reset(p->batch) != MP3B_OK) { t_pool_err = mp3b_last_error(); return MP3B_ERR_CUDA; }
    p->failed = false;
  }
  for (int i = 0; i < p->n; ++i)
    if (!p->slots[i]->open && !p->slots[i]->undelivered) { Slot &s = *p->slots[i]; s.pending = s.done = false; s.flush = s.rc = 0; s.n = 0; s.open = true; p->n_open += 1; *slot = i; return MP3B_OK; }
  return MP3B_ERR_BAD_ARG;                         // every session of the pool is in use
}

static int pool_call(mp3b_pool *p, int slot, const float *pcm, size_t n_floats, int flush, uint8_t *out, size_t cap, size_t *written) {
  if (written) *written = 0;
  if (!p || slot < 0 || slot >= p->n || (n_floats && !pcm)) return MP3B_ERR_BAD_ARG;
  Slot &s = *p->slots[slot];
  // leaves the pool: the last thing a client does with `p` (mp3b_pool_destroy frees it once the count is zero)
  auto leave = [p](int rc) { { std::lock_guard<std::mutex> lk(p->mu); p->in_flight -= 1; p->cv_idle.notify_all(); } return rc; };
  {
    std::unique_lock<std::mutex> lk(p->mu);
    if (s.undelivered) { t_pool_err = "output of the previous call is still pending: call mp3b_pool_take_output"; return MP3B_ERR_BAD_ARG; }
    if (!s.open || s.pending || p->stop) return MP3B_ERR_BAD_ARG;      // one call per session at a time (README:207)
    p->in_flight += 1;
    s.n = n_floats;
    if (n_floats > p->row) s.big.assign(pcm, pcm + n_floats);
    else if (n_floats) memcpy(p->arena + (size_t)slot * p->row, pcm, n_floats * sizeof(float));
    s.flush = flush; s.pending = true;
    { std::lock_guard<std::mutex> g(s.m); s.done = false; }
    if (p->n_pending++ == 0) p->first_pending = std::chrono::steady_clock::now();
    if (p->n_pending == 1 || p->n_pending >= p->n_open) p->cv_worker.notify_one();   // first arrival opens the window, the last closes it
  }
  {
    std::unique_lock<std::mutex> g(s.m);
    s.cv.wait(g, [&] { return s.done; });
    s.done = false;
  }
  {
    std::lock_guard<std::mutex> lk(p->mu);
    if (p->stop && s.pending) {                                       // woken by mp3b_pool_destroy, the request never ran
      s.pending = false;
      t_pool_err = "pool destroyed";
      p->in_flight -= 1; p->cv_idle.notify_all();
      return MP3B_ERR_INTERNAL;
    }
    if (flush) { s.open = false; p->n_open -= 1; p->cv_worker.notify_one(); }   // a flushed session no longer holds steps back
  }
  if (s.rc) { t_pool_err = s.err; return leave(s.rc); }
  if (written) *written = s.out.size();
  if (s.out.size() > cap || (!s.out.empty() && !out)) {               // nothing is lost: the bytes wait for mp3b_pool_take_output
    { std::lock_guard<std::mutex> lk(p->mu); s.undelivered = true; }
    return leave(MP3B_ERR_BUFFER_TOO_SMALL);
  }
  if (!s.out.empty()) memcpy(out, s.out.data(), s.out.size());
  return leave(MP3B_OK);
}

int mp3b_pool_encode(mp3b_pool *p, int slot, const float *pcm, size_t n_floats, uint8_t *out, size_t cap, size_t *written) {
  return pool_call(p, slot, pcm, n_floats, 0, out, cap, written);
}
int mp3b_pool_flush(mp3b_pool *p, int slot, uint8_t *out, size_t cap, size_t *written) {
  return pool_call(p, slot, nullptr, 0, 1, out, cap, written);
}
int mp3b_pool_take_output(mp3b_pool *p, int slot, uint8_t *out, size_t cap, size_t *written) {
  if (!p || slot < 0 || slot >= p->n) return MP3B_ERR_BAD_ARG;
  Slot &s = *p->slots[slot];
  std::lock_guard<std::mutex> lk(p->mu);
  if (!s.undelivered) { if (written) *written = 0; return MP3B_OK; }
  if (written) *written = s.out.size();
  if (s.out.size() > cap || !out) return MP3B_ERR_BUFFER_TOO_SMALL;
  memcpy(out, s.out.data(), s.out.size());
  s.undelivered = false;
  return MP3B_OK;
}
int mp3b_pool_stats(mp3b_pool *p, uint64_t *steps, uint64_t *requests) {
  if (!p) return MP3B_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(p->mu);
  if (steps) *steps = p->steps;
  if (requests) *requests = p->requests;
  return MP3B_OK;
}
const char *mp3b_pool_last_error(void) { return t_pool_err.c_str(); }

}  // extern "C"
