// Session pool: N independent EncoderSessions (SRC:237-350) that are driven from many threads, one blocking
// encode(samples:) / flush() call per session at a time, and advanced on the GPU together.  This is the host-side
// coalescing layer of BASELINE config 5 ("1024 concurrent sessions fed 1152-sample chunks"): the calls that arrive within
// a short window become ONE step of the batch plane — one upload, one pass of the kernels, one download — instead of N
// launch-bound single-stream calls.  A step fires as soon as every session that is currently inside a call has deposited
// its request and either all open sessions are present or `max_wait_us` has elapsed since the first of them arrived.
//
// The worker thread owns the batch (and the CUDA work); client threads only copy their PCM in and their bytes out.
#include <chrono>
#include <condition_variable>
#include <cstring>
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
  int flush = 0, rc = 0;
  std::vector<float> pcm;    // request (copied: the caller's array may be reused after return, SRC:298)
  std::vector<uint8_t> out;  // response
  std::string err;
};
}  // namespace

struct mp3b_pool {
  mp3b_batch *batch = nullptr;
  int n = 0, max_wait_us = 0;
  std::vector<Slot> slots;
  std::mutex mu;
  std::condition_variable cv_worker, cv_client;
  std::thread worker;
  bool stop = false;
  int n_open = 0, n_pending = 0;
  std::chrono::steady_clock::time_point first_pending;
  // statistics
  uint64_t steps = 0, requests = 0;
  // step scratch (worker only)
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
    int any_flush = 0;
    for (int i = 0; i < p->n; ++i) {
      Slot &s = p->slots[i];
      const bool in = s.pending;
      p->ptrs[i] = in && !s.pcm.empty() ? s.pcm.data() : nullptr;
      p->lens[i] = in ? s.pcm.size() : 0;
      p->mask[i] = in && s.flush ? 1 : 0;
      any_flush |= p->mask[i];
      if (in) { p->members.push_back(i); s.pending = false; }
    }
    p->n_pending = 0;
    lk.unlock();
    int rc = mp3b_batch_encode(p->batch, p->ptrs.data(), p->lens.data(), any_flush, any_flush ? p->mask.data() : nullptr);
    std::string err = rc ? mp3b_last_error() : "";
    for (int i : p->members) {
      Slot &s = p->slots[i];                      // only this thread and the (blocked) owner touch a member slot now
      s.rc = rc; s.err = err; s.out.clear();
      if (rc == MP3B_OK) {
        const uint8_t *data = nullptr; size_t len = 0;
        if (mp3b_batch_output(p->batch, i, &data, &len) == MP3B_OK && len) s.out.assign(data, data + len);
        if (s.flush) mp3b_batch_reset_stream(p->batch, i);           // the slot goes back to a fresh EncoderSession
      }
    }
    lk.lock();
    for (int i : p->members) p->slots[i].done = true;
    p->steps += 1; p->requests += p->members.size();
    p->cv_client.notify_all();
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
  p->slots.resize(n_sessions); p->ptrs.resize(n_sessions); p->lens.resize(n_sessions); p->mask.resize(n_sessions);
  p->worker = std::thread(pool_worker, p);
  *out = p;
  return MP3B_OK;
}

void mp3b_pool_destroy(mp3b_pool *p) {
  if (!p) return;
  { std::lock_guard<std::mutex> lk(p->mu); p->stop = true; }
  p->cv_worker.notify_all(); p->cv_client.notify_all();
  if (p->worker.joinable()) p->worker.join();
  mp3b_batch_destroy(p->batch);
  delete p;
}

int mp3b_pool_open(mp3b_pool *p, int *slot) {
  if (!p || !slot) return MP3B_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(p->mu);
  for (int i = 0; i < p->n; ++i)
    if (!p->slots[i].open) { p->slots[i] = Slot(); p->slots[i].open = true; p->n_open += 1; *slot = i; return MP3B_OK; }
  return MP3B_ERR_BAD_ARG;                         // every session of the pool is in use
}

static int pool_call(mp3b_pool *p, int slot, const float *pcm, size_t n_floats, int flush, uint8_t *out, size_t cap, size_t *written) {
  if (written) *written = 0;
  if (!p || slot < 0 || slot >= p->n || (n_floats && !pcm)) return MP3B_ERR_BAD_ARG;
  std::unique_lock<std::mutex> lk(p->mu);
  Slot &s = p->slots[slot];
  if (!s.open || s.pending) return MP3B_ERR_BAD_ARG;                  // one call per session at a time (README:207)
  s.pcm.assign(pcm, pcm + n_floats);
  s.flush = flush; s.done = false; s.pending = true;
  if (p->n_pending++ == 0) p->first_pending = std::chrono::steady_clock::now();
  p->cv_worker.notify_one();
  p->cv_client.wait(lk, [&] { return s.done || p->stop; });
  if (!s.done) return MP3B_ERR_INTERNAL;
  s.done = false;
  if (flush) { s.open = false; p->n_open -= 1; p->cv_worker.notify_one(); }   // a flushed session no longer holds steps back
  if (s.rc) { t_pool_err = s.err; return s.rc; }
  if (written) *written = s.out.size();
  if (s.out.size() > cap) return MP3B_ERR_BUFFER_TOO_SMALL;
  if (!s.out.empty()) memcpy(out, s.out.data(), s.out.size());
  return MP3B_OK;
}

int mp3b_pool_encode(mp3b_pool *p, int slot, const float *pcm, size_t n_floats, uint8_t *out, size_t cap, size_t *written) {
  return pool_call(p, slot, pcm, n_floats, 0, out, cap, written);
}
int mp3b_pool_flush(mp3b_pool *p, int slot, uint8_t *out, size_t cap, size_t *written) {
  return pool_call(p, slot, nullptr, 0, 1, out, cap, written);
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
