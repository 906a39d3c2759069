// topk.cuh -- register-resident sorted top-k list (ascending squared distance) shared by the k-d tree and
// the esrg queries.  KC is the compile-time capacity (all loops fully unrolled, no dynamic indexing, so the
// arrays stay in registers); k <= KC is the run-time length.
//
// The two reference insertion rules differ on ties and both are reproduced:
//   kd_tree.f90:224-236   shift while d2 <  nb(i-1): the new entry goes AFTER equal ones   (BEFORE_EQUAL=false)
//   query_esrg.f90:164-168 position = count(d2_nn < d2)+1: the new entry goes BEFORE equal ones (=true)
// Both insert only if d2 < d(k) (strict), so a tie with the current k-th entry keeps the earlier point.
#pragma once

namespace gi {

template <class R, int KC> struct TopK {
  // The k live entries occupy the TOP k slots, d[KC-k .. KC-1]; the slots below hold -1, which no squared
  // distance sorts before, so they are inert.  That makes the k-th entry the compile-time slot KC-1 for every
  // run-time k: a run-time slot (d[k-1]) would turn the arrays into local memory.
  R d[KC];
  int id[KC];
  R dlast;   // the k-th smallest so far = d[KC-1]
  int base;  // KC - k: slot of the smallest entry

  __device__ __forceinline__ void init(int k, R d_init, int id_init) {
    base = KC - k;
#pragma unroll
    for (int i = 0; i < KC; ++i) { d[i] = i >= base ? d_init : R(-1); id[i] = id_init; }
    dlast = d_init;
  }
  // precondition: d2 < dlast.
  // One predicate per slot, before(j) = "the new entry sorts before d[j]"; d[] is ascending, so before() is
  // monotone in j and slot j either takes its left neighbour (before(j-1)), the new entry (before(j) only) or
  // keeps its value: one compare + predicated moves per slot, no position counting.
  template <bool BEFORE_EQUAL> __device__ __forceinline__ void insert(R d2, int idx) {
    bool below = true;  // before(KC-1) is the precondition
#pragma unroll
    for (int j = KC - 1; j >= 1; --j) {
      const bool lower = BEFORE_EQUAL ? d2 <= d[j - 1] : d2 < d[j - 1];
      d[j] = lower ? d[j - 1] : (below ? d2 : d[j]);  // selects, not branches: ptxas turns the if/else form
      id[j] = lower ? id[j - 1] : (below ? idx : id[j]);  // into 3x as many register moves
      below = lower;
    }
    d[0] = below ? d2 : d[0];
    id[0] = below ? idx : id[0];
    dlast = d[KC - 1];
  }
};

// Any k (the reference takes any num_neighbours, interpolation.f90:107-108,201-202): the same list in global
// scratch memory, element i of a thread's list `stride` elements after element i-1 (coalesced across a warp).
// Same interface and the same two insertion rules; an insertion is a shift loop over memory, so this is the
// slow form -- the kernels instantiate it (KC = 0) only for k > 32.
template <class R> struct TopK<R, 0> {
  R* d;
  int* id;
  int64_t stride;
  int k;
  R dlast;
  static constexpr int base = 0;
  __device__ __forceinline__ void bind(R* d_all, int* id_all, int64_t stride_, int64_t thread) {
    d = d_all + thread; id = id_all + thread; stride = stride_;
  }
  __device__ __forceinline__ void init(int k_, R d_init, int id_init) {
    k = k_;
    for (int i = 0; i < k; ++i) { d[i * stride] = d_init; id[i * stride] = id_init; }
    dlast = d_init;
  }
  __device__ __forceinline__ R dist2(int rank) const { return d[rank * stride]; }
  __device__ __forceinline__ int index(int rank) const { return id[rank * stride]; }
  template <bool BEFORE_EQUAL> __device__ __forceinline__ void insert(R d2, int idx) {  // precondition: d2 < dlast
    int j = k - 1;
    for (; j >= 1; --j) {
      const R prev = d[(j - 1) * stride];
      if (!(BEFORE_EQUAL ? d2 <= prev : d2 < prev)) break;
      d[j * stride] = prev;
      id[j * stride] = id[(j - 1) * stride];
    }
    d[j * stride] = d2;
    id[j * stride] = idx;
    dlast = d[(k - 1) * stride];
  }
};

}  // namespace gi
