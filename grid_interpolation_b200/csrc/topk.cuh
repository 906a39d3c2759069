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
  R d[KC];
  int id[KC];
  R dlast;  // d[k-1]
  int k;

  __device__ __forceinline__ void init(int k_, R d_init, int id_init) {
    k = k_;
#pragma unroll
    for (int i = 0; i < KC; ++i) { d[i] = d_init; id[i] = id_init; }
    dlast = d_init;
  }
  // precondition: d2 < dlast
  template <bool BEFORE_EQUAL> __device__ __forceinline__ void insert(R d2, int idx) {
    int pos = 0;
#pragma unroll
    for (int j = 0; j < KC; ++j) pos += (j < k && (BEFORE_EQUAL ? d[j] < d2 : d[j] <= d2)) ? 1 : 0;
#pragma unroll
    for (int i = KC - 1; i >= 1; --i)
      if (i < k && i > pos) { d[i] = d[i - 1]; id[i] = id[i - 1]; }
#pragma unroll
    for (int i = 0; i < KC; ++i)
      if (i == pos) { d[i] = d2; id[i] = idx; }
#pragma unroll
    for (int i = 0; i < KC; ++i)
      if (i == k - 1) dlast = d[i];
  }
};

}  // namespace gi
