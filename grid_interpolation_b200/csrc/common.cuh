// common.cuh -- shared plumbing of libgridinterp: status codes, stream-ordered scratch memory,
// per-stream device status words, phase timers and the stride-generic array views that let every
// kernel read the reference's Fortran layouts and numpy's C layouts without a transposition pass.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

#include "../../include/gridinterp.h"

namespace gi {

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define GI_CUDA(call)                                                    \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return gi::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
#define GI_TRY(call)                 \
  do {                               \
    int s__ = (call);                \
    if (s__ != GI_OK) return s__;    \
  } while (0)
#define GI_REQUIRE(cond, msg)                   \
  do {                                          \
    if (!(cond)) {                              \
      gi::set_error(std::string("bad argument: ") + msg); \
      return GI_ERR_BAD_ARG;                    \
    }                                           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// kind-dependent constants of the reference (default-kind REAL literals)
// ---------------------------------------------------------------------------------------------
template <class R> struct Consts;
template <> struct Consts<double> {
  __host__ __device__ static double tiny_e4() { return 2.2250738585072014e-308 * 1e4; }  // TINY(1.0)*1e4
  __host__ __device__ static double huge() { return 1.7976931348623157e308; }            // HUGE(1.0)
  __host__ __device__ static double sentinel() { return 1.0e6; }                         // interpolation.f90:159
  __host__ __device__ static double margin() { return -1e-10; }                          // triangle_walk.f90:148
  __host__ __device__ static double ofb() { return -9999.0; }                            // interpolation.f90:44
};
template <> struct Consts<float> {
  __host__ __device__ static float tiny_e4() { return 1.17549435e-38f * 1e4f; }
  __host__ __device__ static float huge() { return 3.40282347e38f; }
  __host__ __device__ static float sentinel() { return 1.0e6f; }
  __host__ __device__ static float margin() { return -1e-10f; }
  __host__ __device__ static float ofb() { return -9999.0f; }
};

// ---------------------------------------------------------------------------------------------
// L2 residency hints (createpolicy + .L2::cache_hint).  For passes that stream large arrays once while gathering
// from arrays that fit the 126 MB L2: the streams are marked evict_first so that they do not push the gathered
// lines out (pack_mesh_kernel: 1.2 GB of index / record traffic against 120 MB of vertices and nodal values).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long l2_evict_last() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_evict_first() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int ldg_hint(const int* p, unsigned long long pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ldg_hint(const double* p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_hint(const float* p, unsigned long long pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ldg_hint(const double2* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ldg_hint(const float2* p, unsigned long long pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}

// ---------------------------------------------------------------------------------------------
// stride-generic views
// ---------------------------------------------------------------------------------------------
template <class R> struct PointsView {  // x(i) = base[i*stride], y(i) = base[i*stride + yoff]
  const R* base;
  int64_t stride, yoff;
  __device__ __forceinline__ R x(int64_t i) const { return __ldg(base + i * stride); }
  __device__ __forceinline__ R y(int64_t i) const { return __ldg(base + i * stride + yoff); }
  // both coordinates of one point: a single 2*sizeof(R) load when the layout is interleaved (and aligned)
  __device__ __forceinline__ void xy(int64_t i, R& px, R& py) const {
    if (stride == 2 && yoff == 1 && (reinterpret_cast<uintptr_t>(base) & (2 * sizeof(R) - 1)) == 0) {
      if (sizeof(R) == 8) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(base) + i);
        px = (R)v.x; py = (R)v.y;
      } else {
        const float2 v = __ldg(reinterpret_cast<const float2*>(base) + i);
        px = (R)v.x; py = (R)v.y;
      }
    } else {
      px = x(i); py = y(i);
    }
  }
  // the same with an L2 cache policy (see l2_evict_last)
  __device__ __forceinline__ void xy(int64_t i, R& px, R& py, unsigned long long pol) const {
    if (stride == 2 && yoff == 1 && (reinterpret_cast<uintptr_t>(base) & (2 * sizeof(R) - 1)) == 0) {
      if (sizeof(R) == 8) {
        const double2 v = ldg_hint(reinterpret_cast<const double2*>(base) + i, pol);
        px = (R)v.x; py = (R)v.y;
      } else {
        const float2 v = ldg_hint(reinterpret_cast<const float2*>(base) + i, pol);
        px = (R)v.x; py = (R)v.y;
      }
    } else {
      px = ldg_hint(base + i * stride, pol); py = ldg_hint(base + i * stride + yoff, pol);
    }
  }
};
// (n,2) array: Fortran order = SoA, C order = interleaved
template <class R> inline PointsView<R> points_n2(const R* p, int64_t n, bool rowmajor) {
  return rowmajor ? PointsView<R>{p, 2, 1} : PointsView<R>{p, 1, n};
}
// (2,n) array (idw_kdtree): Fortran order = interleaved, C order = SoA
template <class R> inline PointsView<R> points_2n(const R* p, int64_t n, bool rowmajor) {
  return rowmajor ? PointsView<R>{p, 1, n} : PointsView<R>{p, 2, 1};
}
struct TopoView {  // a(t,c) = base[t*st + c*sc], t 0-based triangle, c in 0..2
  const int32_t* base;
  int64_t st, sc;
  __device__ __forceinline__ int operator()(int64_t t, int c) const { return __ldg(base + t * st + c * sc); }
  __device__ __forceinline__ int operator()(int64_t t, int c, unsigned long long pol) const {
    return ldg_hint(base + t * st + c * sc, pol);
  }
};
inline TopoView topo_t3(const int32_t* p, int64_t t, bool rowmajor) {
  return rowmajor ? TopoView{p, 3, 1} : TopoView{p, 1, t};
}

// A rectangular window of grid targets addressed by (fast, slow) indices: `fast` runs along the axis
// that is contiguous in the output field.  Row-major field: fast = x; column-major: fast = y.
template <class R> struct TargetWindow {
  const R* fast_axis;  // device pointer, full axis
  const R* slow_axis;
  int nf_full, ns_full;  // full grid extents along fast / slow
  int f0, f1, s0, s1;    // window [f0,f1) x [s0,s1)
  int fast_is_y;         // 1: fast axis = y (column-major field), 0: fast axis = x
  __host__ __device__ int nf() const { return f1 - f0; }
  __host__ __device__ int ns() const { return s1 - s0; }
};

// ---------------------------------------------------------------------------------------------
// stream-ordered scratch memory (cudaMallocAsync pool, cached between calls)
// ---------------------------------------------------------------------------------------------
int ensure_device();  // validates that a CUDA device exists; configures the mem pool once per device

// Small allocations come from a per-(device, stream) arena: one cached block that the Scratch objects alive on that
// stream bump-allocate from and that is rewound when the last of them is released.  A call at the reference's own
// test size (5 000 points -> 132 x 79 targets) makes ~15 allocations of a few hundred KB in total; as
// cudaMallocAsync / cudaFreeAsync pairs they cost more host time than its kernels take.  Everything that touches
// arena memory is ordered on the arena's stream, so the next call may reuse it without waiting.
struct ScratchArena;
ScratchArena* arena_acquire(cudaStream_t s);            // nullptr: no arena (allocation failed)
void* arena_alloc(ScratchArena* a, size_t bytes);       // nullptr: does not fit
void arena_release(ScratchArena* a);
void arena_trim();                                      // frees the idle arenas of the current device
constexpr size_t kArenaMaxAlloc = (size_t)2 << 20;      // larger requests go to the stream-ordered pool

class Scratch {
 public:
  // persistent = true: plain cudaMalloc / cudaFree, for memory that outlives the call and its stream (plans)
  explicit Scratch(cudaStream_t s, bool persistent = false) : stream_(s), persistent_(persistent) {}
  ~Scratch() { release(); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  template <class T> int alloc(T** out, size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    if (!persistent_ && bytes <= kArenaMaxAlloc) {
      if (!arena_ && !arena_tried_) { arena_ = arena_acquire(stream_); arena_tried_ = true; }
      if (arena_ && (p = arena_alloc(arena_, bytes))) {
        *out = static_cast<T*>(p);
        return GI_OK;
      }
    }
    cudaError_t e = persistent_ ? cudaMalloc(&p, bytes) : cudaMallocAsync(&p, bytes, stream_);
    if (e != cudaSuccess) {
      *out = nullptr;
      if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        set_error("out of device memory allocating " + std::to_string(bytes) + " bytes");
        return GI_ERR_OOM;
      }
      return cuda_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
    }
    ptrs_.push_back(p);
    *out = static_cast<T*>(p);
    return GI_OK;
  }
  void release() {
    for (void* p : ptrs_) {
      if (persistent_) cudaFree(p);
      else cudaFreeAsync(p, stream_);
    }
    ptrs_.clear();
    if (arena_) { arena_release(arena_); arena_ = nullptr; }
    arena_tried_ = false;
  }
  cudaStream_t stream() const { return stream_; }
  void set_stream(cudaStream_t s) { stream_ = s; }   // persistent scratch: the stream its owner's kernels run on

 private:
  cudaStream_t stream_;
  bool persistent_;
  std::vector<void*> ptrs_;
  ScratchArena* arena_ = nullptr;
  bool arena_tried_ = false;
};

// ---------------------------------------------------------------------------------------------
// per-stream status word: kernels raise it with atomicMax, flush_status() forwards it to a pinned
// host slot at the end of every call; gi_stream_status() synchronises and reads the slot.
// ---------------------------------------------------------------------------------------------
struct StatusSlot {
  int* dev;        // device word
  int* host;       // pinned, mapped
  int* host_dev;   // device alias of host
};
int get_status_slot(cudaStream_t s, StatusSlot* out);
int flush_status(cudaStream_t s, const StatusSlot& slot);
int read_status(cudaStream_t s, bool synchronize);

__device__ __forceinline__ void raise_status(int* status, int code) { atomicMax(status, code); }

// ---------------------------------------------------------------------------------------------
// phase timers (CUDA events on the launching stream)
// ---------------------------------------------------------------------------------------------
class Phases {
 public:
  explicit Phases(cudaStream_t s);
  ~Phases() { finish(); }
  void mark(const char* name);  // closes the previous phase, opens `name`
  void finish();                // closes the last phase; results are resolved lazily
 private:
  cudaStream_t stream_;
  bool on_;
};
void set_profiling(bool on);
bool profiling();

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// GI_HOST calls: staging helpers and the copy / compute pipeline.
//
// A host call is bound by PCIe, not by the kernels (C4: 0.36 GB in, 2.1 GB out, 3.4 ms of compute), so the grid
// operators cut their output into bands along the axis that is contiguous in memory and copy band b back on a
// second stream while band b+1 is being computed; bilinear does the same with chunks of points in both
// directions.  Results are identical to the one-shot path: every target is computed by the same kernels.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxBands = 16;
struct HostPipe {
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t computed[kMaxBands] = {};  // band b is ready on the device
  cudaEvent_t arrived[kMaxBands] = {};   // chunk b of the inputs is on the device
};
int host_pipe(HostPipe** out);  // per host thread, created on first use (api.cu)
// bands of at least ~32 MB of output (a PCIe transfer of that size is near its peak rate), at most kMaxBands
int64_t pipeline_band_bytes();  // gi_set_pipeline_band_bytes(), default 32 MB (api.cu)
int64_t fixup_list_capacity();  // gi_set_fixup_list_capacity(), default 4 Mi targets (api.cu)
double sliver_guard_quality();  // gi_set_sliver_guard(), default 0 = off (api.cu, barycentric.cu sliver_quality)
inline int plan_bands(int64_t extent, int64_t bytes) {
  int64_t nb = bytes / pipeline_band_bytes();
  if (nb > kMaxBands) nb = kMaxBands;
  if (nb > extent) nb = extent;
  return nb < 1 ? 1 : (int)nb;
}
inline void band_range(int64_t lo, int64_t hi, int nb, int b, int64_t align, int64_t* b0, int64_t* b1) {
  const int64_t n = hi - lo;
  int64_t per = (n + nb - 1) / nb;
  per = (per + align - 1) / align * align;
  *b0 = lo + per * b < hi ? lo + per * b : hi;
  *b1 = *b0 + per < hi ? *b0 + per : hi;
}

// A plan (gi_plan): search structure / packed mesh + target grid resident on the device; execute() interpolates
// one field of nodal values.  One implementation per operator (kdtree.cu, esrg.cu, barycentric.cu).
struct PlanBase {
  int real_bytes = 8;  // sizeof(R) of the plan
  int device = 0;
  virtual ~PlanBase() {}
  // data_in (num_points) -> data_ip (band rows, len_x) in the layout fixed at creation; mem / stream as in the
  // one-shot calls
  virtual int execute(const void* data_in, void* data_ip, int mem, cudaStream_t st) = 0;
};

template <class T> int to_device(Scratch& scr, const T* host, size_t count, T** dev) {
  GI_TRY(scr.alloc(dev, count));
  if (count) GI_CUDA(cudaMemcpyAsync(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice, scr.stream()));
  return GI_OK;
}
template <class T> int to_host(cudaStream_t st, T* host, const T* dev, size_t count) {
  if (host && count) GI_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st));
  return GI_OK;
}
template <class T> int out_device(Scratch& scr, T* host, size_t count, T** dev) {  // device twin of an optional host output
  *dev = nullptr;
  if (host) GI_TRY(scr.alloc(dev, count));
  return GI_OK;
}

}  // namespace gi
