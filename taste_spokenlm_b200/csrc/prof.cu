// Launch accounting and optional per-kernel-class device timing (CUDA events on the launching stream).
// bench.py reads this for `gpu_launches` and for the live `roofline` figures; it is off by default and costs one
// atomic increment per launch when off.
#include <atomic>
#include <mutex>
#include <vector>

#include "internal.h"

namespace taste {

static const char* kClassNames[KC_COUNT] = {
    "gemm_bf16_tcgen05", "attention_encoder", "attention_aggregator", "layernorm", "cast_bf16", "logmel_tile",
    "logmel_finish",     "embed",             "word_pool",            "rvq_encode", "rvq_decode", "map_llm",
    "attention_tcgen05", "resample_mean",     "logmel_split",         "logmel_dft", "logmel_mel", "rvq_project",
};

struct ProfRecord {
  int kc;
  cudaEvent_t e0, e1;
  double flops, bytes;
};

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_enabled{0};
static std::mutex g_mu;
static std::vector<ProfRecord> g_records;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(cudaStream_t stream, int kc, double flops, double bytes) : stream_(stream), slot_(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_enabled.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  ProfRecord r;
  r.kc = kc;
  r.e0 = take_event();
  r.e1 = take_event();
  r.flops = flops;
  r.bytes = bytes;
  cudaEventRecord(r.e0, stream);
  g_records.push_back(r);
  slot_ = int(g_records.size()) - 1;
}

ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot_ < int(g_records.size())) cudaEventRecord(g_records[slot_].e1, stream_);
}

}  // namespace taste

using namespace taste;

extern "C" {

unsigned long long taste_launch_count(void) { return g_launches.load(); }

int taste_prof_enable(int on) {
  g_enabled.store(on ? 1 : 0);
  return 0;
}

int taste_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_records) {
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  g_records.clear();
  return 0;
}

int taste_prof_collect(taste_prof_entry_t* out, int max_entries, int* n_out) {
  if (!out || !n_out || max_entries <= 0) return set_error(TASTE_E_ARG, "prof_collect: null output");
  std::lock_guard<std::mutex> lk(g_mu);
  taste_prof_entry_t acc[KC_COUNT];
  for (int i = 0; i < KC_COUNT; ++i) {
    acc[i].name = kClassNames[i];
    acc[i].launches = 0;
    acc[i].total_ms = acc[i].flops = acc[i].bytes = 0.0;
  }
  for (auto& r : g_records) {
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e != cudaSuccess) return set_error((int)e, "prof_collect: %s", cudaGetErrorString(e));
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (e != cudaSuccess) return set_error((int)e, "prof_collect: %s", cudaGetErrorString(e));
    acc[r.kc].launches += 1;
    acc[r.kc].total_ms += ms;
    acc[r.kc].flops += r.flops;
    acc[r.kc].bytes += r.bytes;
  }
  int n = 0;
  for (int i = 0; i < KC_COUNT && n < max_entries; ++i)
    if (acc[i].launches > 0) out[n++] = acc[i];
  *n_out = n;
  return 0;
}

}  // extern "C"
