// Executor-level CUDA graph cache shared by the encoder and decoder executors (swin_exec.cu, fpn_exec.cu).
#pragma once
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace mtus_graphs {

// ---- executor-level CUDA graph cache ------------------------------------------------------------------------------
// One executor call is 170-330 launches whose arguments are fully determined by (config, pointers, ranges).  PyTorch's
// caching allocator hands the same blocks back step after step, so the schedule is captured once per distinct
// argument set (stream capture, the side-stream fork / join included) and replayed with one cudaGraphLaunch: the host
// cost of a call drops from milliseconds to microseconds and kernel-to-kernel gaps shrink.  Keys are exact (every
// pointer / scalar that reaches a kernel); a miss costs one capture + instantiate.  MTUS_GRAPHS=0 disables the cache;
// calls made while the stream is already being captured by the caller run the plain schedule.
struct GraphEntry { std::vector<uint8_t> key; cudaGraphExec_t exec; int64_t launches; uint64_t stamp; };
inline std::vector<GraphEntry> g_graphs;
inline uint64_t g_stamp = 0;
inline int64_t g_graph_hits = 0, g_graph_misses = 0;
constexpr size_t kMaxGraphs = 48;

inline bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MTUS_GRAPHS");
    v = (e && atoi(e) == 0) ? 0 : 1;
    const char* t = getenv("MTUS_TIME_KERNELS");
    if (t && atoi(t) != 0) v = 0;
    // a kernel profiler (ncu / nsys injection) wants to see kernel launches, not graph launches, and replays each kernel:
    // run the plain schedule under it unless MTUS_GRAPHS=1 insists
    if (!e) {
      static const char* inj[] = {"CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "NSYS_PROFILING_SESSION_ID"};
      for (const char* n : inj) { const char* x = getenv(n); if (x && *x) v = 0; }
    }
  }
  return v == 1;
}

// per configuration (the key minus its address fields): how often its graphs were reused / had to be rebuilt
struct ShapeStat { std::vector<uint8_t> shape; int64_t hits, misses; };
inline std::vector<ShapeStat> g_shapes;

struct KeyBuilder {
  std::vector<uint8_t> k;
  size_t shape_len = 0;                    // leading bytes that describe the configuration, not addresses
  KeyBuilder& end_shape() { shape_len = k.size(); return *this; }
  template <typename T> KeyBuilder& add(const T& v) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
    k.insert(k.end(), p, p + sizeof(T));
    return *this;
  }
};

// body(stream) enqueues the schedule on `stream`.  Capture always happens on an internal stream (the caller's stream is
// usually PyTorch's legacy default stream, which cannot be captured); the instantiated graph is then launched into the
// caller's stream, which orders it like any other work there.
inline cudaStream_t capture_stream() {
  static cudaStream_t s[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!s[dev] && cudaStreamCreateWithFlags(&s[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return s[dev];
}

inline ShapeStat& shape_stat(const std::vector<uint8_t>& key, size_t shape_len) {
  if (shape_len == 0 || shape_len > key.size()) shape_len = key.size();
  for (ShapeStat& s : g_shapes)
    if (s.shape.size() == shape_len && memcmp(s.shape.data(), key.data(), shape_len) == 0) return s;
  if (g_shapes.size() >= 256) g_shapes.clear();
  g_shapes.push_back(ShapeStat{std::vector<uint8_t>(key.begin(), key.begin() + shape_len), 0, 0});
  return g_shapes.back();
}

template <typename F>
int run_cached(const std::vector<uint8_t>& key, cudaStream_t st, F&& body, size_t shape_len = 0) {
  if (!graphs_enabled()) return body((void*)st);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return body((void*)st); }
  for (GraphEntry& e : g_graphs) {
    if (e.key == key) {
      e.stamp = ++g_stamp; ++g_graph_hits; ++shape_stat(key, shape_len).hits;
      cudaError_t le = cudaGraphLaunch(e.exec, st);
      if (le != cudaSuccess) return (int)le;
      mtus_internal_count_launches((int)e.launches);
      return MTUS_OK;
    }
  }
  ++g_graph_misses;
  // addresses that keep changing would mean one capture + instantiate per call: stop adding graphs for a configuration
  // whose graphs clearly are not reused (lookups of the graphs already built continue)
  {
    ShapeStat& ss = shape_stat(key, shape_len);
    ++ss.misses;
    if (ss.misses > 8 && ss.hits < 4 * ss.misses) return body((void*)st);
  }
  cudaStream_t cap = capture_stream();
  if (!cap || cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return body((void*)st); }
  const int64_t before = mtus_launch_count();
  const int rc = body((void*)cap);
  const int64_t launches = mtus_launch_count() - before;
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(cap, &graph);
  mtus_internal_count_launches(-(int)launches);             // nothing ran yet: counted when the graph (or the eager retry) runs
  if (rc != MTUS_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
  if (ce != cudaSuccess || !graph) {
    if (getenv("MTUS_GRAPH_DEBUG")) fprintf(stderr, "mtus graph cache: capture failed (%s), running the plain schedule\n", cudaGetErrorString(ce));
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return body((void*)st);
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess || !exec) {
    if (getenv("MTUS_GRAPH_DEBUG")) fprintf(stderr, "mtus graph cache: instantiate failed (%s), running the plain schedule\n", cudaGetErrorString(ce));
    cudaGetLastError();
    return body((void*)st);
  }
  if (g_graphs.size() >= kMaxGraphs) {                      // evict the least recently used entry
    size_t lru = 0;
    for (size_t i = 1; i < g_graphs.size(); ++i) if (g_graphs[i].stamp < g_graphs[lru].stamp) lru = i;
    cudaGraphExecDestroy(g_graphs[lru].exec);
    g_graphs.erase(g_graphs.begin() + lru);
  }
  g_graphs.push_back(GraphEntry{key, exec, launches, ++g_stamp});
  ce = cudaGraphLaunch(exec, st);
  if (ce != cudaSuccess) return (int)ce;
  mtus_internal_count_launches((int)launches);
  return MTUS_OK;
}


template <typename F>
int run_cached(const KeyBuilder& kb, cudaStream_t st, F&& body) { return run_cached(kb.k, st, body, kb.shape_len); }

}  // namespace mtus_graphs
