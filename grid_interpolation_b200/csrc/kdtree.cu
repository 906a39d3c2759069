// kdtree.cu -- IDW from scattered points to a regular grid with the k-d tree k-nearest-neighbour search.
//
// Replaces (reference file:line):
//   interpolation.f90:107-194   idw_kdtree (operator + IDW weights)
//   kd_tree.f90:30-155          build_kd_tree / build_kd_tree_sub / partition_iter
//   kd_tree.f90:161-251         nearest_neighbours / insert_neighbour / distance_squared
//
// B200 design (not a port):
//   build   The reference's tree is canonical for distinct coordinates (node = element of rank (n+1)/2
//           along axis depth mod 2, kd_tree.f90:45,54-68), so it is built here without any
//           recursion or pointers: both coordinate orders are sorted once (hand-written stable LSD radix
//           sort, 8 passes of 8 bits) and the tree is then produced level by level -- in
//           every segment the axis-sorted list already has the median in place, and the other list is
//           stably partitioned around it with one device-wide prefix sum per level.  The result is an
//           IMPLICIT tree: the node of segment [lo,hi) is element lo+(hi-lo+1)/2-1 of one array, its
//           subtrees are the two sub-segments.  With tied coordinates the reference's tree depends on its
//           quickselect's swap history instead: the build detects ties and evaluates the reference's Lomuto
//           passes in parallel (kd_build_faithful below) -- same tree, node by node.
//   query   one thread per grid target, explicit stack of deferred far branches (depth <= 32),
//           register-resident sorted top-k (topk.cuh).  The far branch is taken iff
//           abs(diff) < max d2 -- the reference's rule, which compares a distance with a squared distance
//           (kd_tree.f90:191,198) and must be reproduced for index parity -- AND diff*diff < max d2.  The
//           second condition only skips subtrees in which every point has d2 >= diff*diff >= max d2, i.e.
//           none could be inserted (insertion needs d2 < max d2, which only shrinks), so the neighbour
//           list after the skipped visit is bit-identical to the reference's; it cuts node visits ~14x when
//           the k-th d2 is > 1.  GI_KD_REFERENCE_VISITS turns it off (visit-count parity with the oracle).
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <memory>

#include "kernels.cuh"
#include "radix.cuh"
#include "scan.cuh"
#include "topk.cuh"

namespace gi {
namespace {
namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------
// sort keys: both coordinate orders are produced by radix.cuh's stable sort on an order-preserving 64-bit image
// of the coordinate (ties end up in id order).  The first version used a bitonic network: 56 launches per sort,
// 0.65 ms for 1 M points, launch-bound.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

template <class R>
__global__ void __launch_bounds__(256)
sort_keys_kernel(PointsView<R> pts, int n, int axis, unsigned long long* __restrict__ keys, int* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = order_key((double)(axis == 0 ? pts.x(i) : pts.y(i)));
  ids[i] = i;
}

// ---------------------------------------------------------------------------------------------
// level-synchronous canonical build
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kd_init_kernel(const int* __restrict__ list_x, const int* __restrict__ list_y, int n, int* __restrict__ pos_x,
               int* __restrict__ pos_y, int* __restrict__ seg_lo, int* __restrict__ seg_hi) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  pos_x[list_x[p]] = p;
  pos_y[list_y[p]] = p;
  seg_lo[p] = 0;
  seg_hi[p] = n;
}

// flag[p] = 1 iff the point at position p of the OTHER list belongs to the left part of its segment
__global__ void __launch_bounds__(256)
kd_flag_kernel(const int* __restrict__ other, const int* __restrict__ pos_split, const int* __restrict__ seg_lo,
               const int* __restrict__ seg_hi, int n, int* __restrict__ flag) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int lo = seg_lo[p];
  int f = 0;
  if (lo >= 0) {
    const int hi = seg_hi[p];
    const int m = lo + ((hi - lo + 1) >> 1) - 1;  // ind_median = (num_points + 1) / 2   (kd_tree.f90:54)
    f = pos_split[other[p]] < m ? 1 : 0;
  }
  flag[p] = f;
}

__global__ void __launch_bounds__(256)
kd_partition_kernel(const int* __restrict__ split, const int* __restrict__ other, const int* __restrict__ pos_split,
                    const int* __restrict__ pos_other, const int* __restrict__ seg_lo,
                    const int* __restrict__ seg_hi, const int* __restrict__ scan, int n,
                    int* __restrict__ other_new, int* __restrict__ pos_other_new, int* __restrict__ seg_lo_new,
                    int* __restrict__ seg_hi_new) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int lo = seg_lo[p];
  const int pt = other[p];
  if (lo < 0) {  // position already holds a finished node
    other_new[p] = pt;
    pos_other_new[pt] = p;
    seg_lo_new[p] = -1;
    seg_hi_new[p] = -1;
    return;
  }
  const int hi = seg_hi[p];
  const int m = lo + ((hi - lo + 1) >> 1) - 1;
  const int ps = pos_split[pt];
  const int left_before = scan[p] - scan[lo];
  int newp;
  if (ps < m) newp = lo + left_before;                       // left subtree keeps its relative order
  else if (ps == m) newp = m;                                // the node
  else {
    const int node_pos = pos_other[split[m]];
    newp = m + 1 + ((p - lo) - left_before - (node_pos < p ? 1 : 0));
  }
  other_new[newp] = pt;
  pos_other_new[pt] = newp;
  // the segment bookkeeping is per POSITION
  if (p < m) { seg_lo_new[p] = lo; seg_hi_new[p] = m; }
  else if (p == m) { seg_lo_new[p] = -1; seg_hi_new[p] = -1; }
  else { seg_lo_new[p] = m + 1; seg_hi_new[p] = hi; }
}

// The last levels in shared memory: once every segment has at most kFinishMax points (n >> d0 <= kFinishMax), one
// CTA per segment runs the remaining levels of the same algorithm -- flag, prefix sum, stable partition of the other
// list around the median of the split list -- on local positions, with barriers instead of launches: 1 launch
// instead of 5 per level for the lower half of the tree (1 M points: levels 10..19, 0.45 of the build's 1.35 ms;
// 5 000 points: 2 levels by launches instead of 13).  CTA b finds its segment by descending d0 levels along the bits
// of b -- the segment bounds are pure arithmetic (kd_tree.f90:54).
constexpr int kFinishMax = 1024;
constexpr int kFinishItems = kFinishMax / 256;
__global__ void __launch_bounds__(256)
kd_finish_kernel(int* __restrict__ list_x, const int* __restrict__ list_y, const int* __restrict__ pos_x,
                 const int* __restrict__ pos_y, int n, int d0, int levels) {
  __shared__ int L[3][kFinishMax];     // point ids by position: two lists + the partition's destination
  __shared__ short X[3][kFinishMax];   // X[a][i]: position in the OTHER list of the point at position i of list a
  __shared__ short slo[kFinishMax], shi[kFinishMax];
  __shared__ int scan[kFinishMax];
  int lo = 0, hi = n;
  for (int j = d0 - 1; j >= 0 && lo < hi; --j) {
    const int m = lo + ((hi - lo + 1) >> 1) - 1;   // ind_median = (num_points + 1) / 2   (kd_tree.f90:54)
    if ((blockIdx.x >> j) & 1) lo = m + 1; else hi = m;
  }
  const int M = hi - lo;
  if (M <= 0) return;
  const int tid = threadIdx.x;
  for (int i = tid; i < M; i += 256) {
    const int px = list_x[lo + i], py = list_y[lo + i];
    L[0][i] = px; L[1][i] = py;
    X[0][i] = (short)(pos_y[px] - lo);
    X[1][i] = (short)(pos_x[py] - lo);
    slo[i] = 0; shi[i] = (short)M;
  }
  int buf[2] = {0, 1}, spare = 2;   // which of L / X holds list x, list y
  __syncthreads();
  for (int d = d0; d < levels; ++d) {
    const int ax = d & 1, ot = ax ^ 1;  // axis = mod(depth, 2)   (kd_tree.f90:45)
    const int bs = buf[ax], bo = buf[ot];
    int l[kFinishItems], m[kFinishItems], f[kFinishItems];
    int acc = 0;
#pragma unroll
    for (int i = 0; i < kFinishItems; ++i) {
      const int p = tid * kFinishItems + i;
      l[i] = -1; m[i] = 0; f[i] = 0;
      if (p < M) {
        l[i] = slo[p];
        if (l[i] >= 0) {
          const int h = shi[p];
          m[i] = l[i] + ((h - l[i] + 1) >> 1) - 1;
          f[i] = X[bo][p] < m[i] ? 1 : 0;      // the point at position p of the other list goes left
        }
      }
      acc += f[i];
    }
    int total;
    int run = block_exclusive_scan_256(acc, &total);
#pragma unroll
    for (int i = 0; i < kFinishItems; ++i) {
      const int p = tid * kFinishItems + i;
      if (p < M) scan[p] = run;
      run += f[i];
    }
    __syncthreads();
    int newp[kFinishItems], ps[kFinishItems];
#pragma unroll
    for (int i = 0; i < kFinishItems; ++i) {
      const int p = tid * kFinishItems + i;
      newp[i] = p; ps[i] = 0;
      if (p < M) {
        ps[i] = X[bo][p];
        if (l[i] >= 0) {
          const int left_before = scan[p] - scan[l[i]];
          if (ps[i] < m[i]) newp[i] = l[i] + left_before;          // left subtree keeps its relative order
          else if (ps[i] == m[i]) newp[i] = m[i];                  // the node
          else {
            const int node_pos = X[bs][m[i]];
            newp[i] = m[i] + 1 + ((p - l[i]) - left_before - (node_pos < p ? 1 : 0));
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kFinishItems; ++i) {
      const int p = tid * kFinishItems + i;
      if (p < M) {
        L[spare][newp[i]] = L[bo][p];
        X[spare][newp[i]] = (short)ps[i];
        if (l[i] >= 0) {
          X[bs][ps[i]] = (short)newp[i];
          // the segment bookkeeping is per POSITION
          const int h = shi[p];
          if (p < m[i]) { shi[p] = (short)m[i]; }
          else if (p == m[i]) { slo[p] = -1; shi[p] = -1; }
          else { slo[p] = (short)(m[i] + 1); shi[p] = (short)h; }
        }
      }
    }
    __syncthreads();
    buf[ot] = spare;
    spare = bo;
  }
  for (int i = tid; i < M; i += 256) list_x[lo + i] = L[buf[0]][i];   // both lists are the tree order now
}

template <class R> struct alignas(16) XY { R x, y; };

template <class R>
__global__ void __launch_bounds__(256)
kd_gather_kernel(PointsView<R> pts, const R* __restrict__ data, const int* __restrict__ order, int n,
                 XY<R>* __restrict__ nodes, R* __restrict__ nval, int* __restrict__ order_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = order[p];
  if (nodes) { XY<R> v; v.x = pts.x(i); v.y = pts.y(i); nodes[p] = v; }
  if (nval) nval[p] = __ldg(data + i);
  if (order_out) order_out[p] = i + 1;
}

// ---------------------------------------------------------------------------------------------
// Coordinate ties: the quickselect-faithful build.
//
// With tied coordinates the reference's tree is no longer determined by ranks: equal keys stay right of a Lomuto
// pivot (kd_tree.f90:129), which tied element lands on the median and in which order a child receives its
// sub-array depend on the swap history of every ancestor's quickselect (kd_tree.f90:97-155).  The tree must
// then be built the way the reference builds it -- but a Lomuto pass is a deterministic permutation that can be
// evaluated in parallel:
//   * the elements < pivot end up at the front in encounter order (stable);
//   * the elements >= pivot form a FIFO queue in positions [store, i): a later ">=" element is appended at the
//     back, and every "<" element met after the first ">=" one (position i0) moves the queue's front to its back
//     (the swap at :131-136).  From i0 on every loop iteration therefore appends exactly one item; item number s
//     is appended by iteration e = i0 + s and is either the element found there, or -- if that element was "<" --
//     the item popped by that swap, which is item number t = #("<" in [i0, e)) < s.  The queue finally holds items
//     T .. P-1 (T = number of pops) in positions store .. hi-1, before the pivot is swapped in at `store` (:141-146).
//   So the source of every destination follows from one prefix sum F over the "<" flags: walk s -> F[e] - F[i0]
//   until an iteration with a ">=" element is reached (one hop per time the reference swapped that item).
// Every unfinished node performs one pass per round (its own [ind_left, ind_right] range, its own axis), nodes
// whose quickselect has converged hand their two sub-ranges to their children (:64-68), all in the same
// launches; the array the rounds leave behind is the implicit tree the query reads.  Rounds ~ sum over levels of
// the passes per quickselect (a few hundred for 1e6 points): ~10x slower than the rank-based build, used only
// when a tie is detected (or GI_KD_FAITHFUL_BUILD asks for it).
// ---------------------------------------------------------------------------------------------
struct alignas(16) KdfNode {
  int a_lo, a_hi;  // the quickselect's current [ind_left, ind_right] (0-based, inclusive)
  int n_hi;        // last position of the node's sub-array (its first position is the node's id)
  int depth;
};

// Sub-arrays of at most kSerialMax points do not take part in the rounds: a round costs ~100 us whatever the size of
// the ranges, and half of the few hundred rounds of a 1 M-point build belong to the lowest ten levels.  One warp per
// such sub-array copies it to shared memory and lane 0 runs the reference's recursion on it as written
// (build_kd_tree / partition_iter, kd_tree.f90:30-155, in place): a thousand serial builds side by side, ~1.4 ms.
constexpr int kSerialMax = 1024;
struct KdfSmall { int lo, hi, depth; };   // positions [lo, hi] of the array, depth of the sub-tree's root

template <class R>
__global__ void __launch_bounds__(32)
kdf_serial_kernel(R* __restrict__ X, R* __restrict__ Y, int* __restrict__ ID, const KdfSmall* __restrict__ work,
                  const int* __restrict__ count) {
  __shared__ R sx[kSerialMax], sy[kSerialMax];
  __shared__ int sid[kSerialMax];
  const int lane = threadIdx.x;
  const int nwork = *count;
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const KdfSmall it = work[wi];
    const int m = it.hi - it.lo + 1;
    for (int i = lane; i < m; i += 32) { sx[i] = X[it.lo + i]; sy[i] = Y[it.lo + i]; sid[i] = ID[it.lo + i]; }
    __syncwarp();
    if (lane == 0) {
      auto swap3 = [&](int a, int b) {
        const R tx = sx[a], ty = sy[a]; const int ti = sid[a];
        sx[a] = sx[b]; sy[a] = sy[b]; sid[a] = sid[b];
        sx[b] = tx; sy[b] = ty; sid[b] = ti;
      };
      int s_lo[32], s_hi[32], s_d[32];   // sub-arrays still to be built (the recursion of :64-68)
      int sp = 1;
      s_lo[0] = 0; s_hi[0] = m - 1; s_d[0] = it.depth;
      while (sp > 0) {
        --sp;
        const int lo = s_lo[sp], hi = s_hi[sp], d = s_d[sp];
        const int k = lo + ((hi - lo + 2) >> 1) - 1;           // ind_median = (num_points + 1) / 2   (:54)
        const R* K = (d & 1) ? sy : sx;                        // axis = mod(depth, 2)   (:45)
        int a = lo, b = hi;
        while (a != b) {                                       // partition_iter (:112)
          const int pidx = (a + b) >> 1;                       // :115
          const R pivot = K[pidx];
          swap3(pidx, b);                                      // :119-124
          int store = a;
          for (int i = a; i < b; ++i) {                        // :128-138
            if (K[i] < pivot) { swap3(store, i); ++store; }
          }
          swap3(store, b);                                     // :141-146
          if (k < store) b = store - 1;                        // :149-153
          else if (k > store) a = store + 1;
          else break;
        }
        if (k > lo && sp < 32) { s_lo[sp] = lo; s_hi[sp] = k - 1; s_d[sp] = d + 1; ++sp; }
        if (k < hi && sp < 32) { s_lo[sp] = k + 1; s_hi[sp] = hi; s_d[sp] = d + 1; ++sp; }
      }
    }
    __syncwarp();
    for (int i = lane; i < m; i += 32) { X[it.lo + i] = sx[i]; Y[it.lo + i] = sy[i]; ID[it.lo + i] = sid[i]; }
    __syncwarp();
  }
}

// position i of the array as the pass sees it, i.e. after pivot and last element were exchanged (:119-124)
__device__ __forceinline__ int kdf_virt(int i, int pidx, int hi) { return i == pidx ? hi : (i == hi ? pidx : i); }

template <class R>
__global__ void __launch_bounds__(256)
kdf_init_kernel(PointsView<R> pts, int n, R* __restrict__ X, R* __restrict__ Y, int* __restrict__ ID,
                int* __restrict__ nid, KdfNode* __restrict__ nodes, int* __restrict__ first_ge) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  X[p] = pts.x(p); Y[p] = pts.y(p); ID[p] = p;
  nid[p] = 0;
  if (p == 0) { nodes[0] = KdfNode{0, n - 1, n - 1, 0}; first_ge[0] = n - 1; }
}

template <class R>
__device__ __forceinline__ int kdf_flag_at(int p, const R* __restrict__ X, const R* __restrict__ Y,
                                           const int* __restrict__ nid, const KdfNode* __restrict__ nodes) {
  int f = 0;
  const int node = nid[p];
  if (node >= 0) {
    const KdfNode nd = nodes[node];
    if (p >= nd.a_lo && p < nd.a_hi) {
      const R* K = (nd.depth & 1) ? Y : X;                 // axis = mod(depth, 2)   (kd_tree.f90:45)
      const int pidx = (nd.a_lo + nd.a_hi) >> 1;           // :115
      f = K[p == pidx ? nd.a_hi : p] < K[pidx] ? 1 : 0;    // :129 (strict)
    }
  }
  return f;
}
template <class R>
__global__ void __launch_bounds__(256)
kdf_flag_kernel(const R* __restrict__ X, const R* __restrict__ Y, const int* __restrict__ nid,
                const KdfNode* __restrict__ nodes, int n, int* __restrict__ flag) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  flag[p] = kdf_flag_at<R>(p, X, Y, nid, nodes);
}

// per node: i0 = first position of the pass whose element is not "<" (a_hi if there is none)
__device__ __forceinline__ void kdf_first_ge_at(int p, const int* __restrict__ nid, const KdfNode* __restrict__ nodes,
                                                const int* __restrict__ F, int* __restrict__ first_ge) {
  const int node = nid[p];
  if (node < 0) return;
  const KdfNode nd = nodes[node];
  if (p != nd.a_hi || nd.a_lo == nd.a_hi) return;
  const int base = F[nd.a_lo] - nd.a_lo;
  int l = nd.a_lo, r = nd.a_hi;  // largest i with "all of [a_lo, i) are <"
  while (l < r) {
    const int mid = (l + r + 1) >> 1;
    if (F[mid] - mid == base) l = mid; else r = mid - 1;
  }
  first_ge[node] = l;
}
__global__ void __launch_bounds__(256)
kdf_first_ge_kernel(const int* __restrict__ nid, const KdfNode* __restrict__ nodes, const int* __restrict__ F, int n,
                    int* __restrict__ first_ge) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  kdf_first_ge_at(p, nid, nodes, F, first_ge);
}

template <class R>
__device__ __forceinline__ void
kdf_pass_at(int p, const R* __restrict__ X, const R* __restrict__ Y, const int* __restrict__ ID, int* __restrict__ nid,
            const KdfNode* __restrict__ nodes, const int* __restrict__ flag, const int* __restrict__ F,
            const int* __restrict__ first_ge, R* __restrict__ X2, R* __restrict__ Y2,
            int* __restrict__ ID2, KdfNode* __restrict__ nodes_next, int* __restrict__ alive,
            KdfSmall* __restrict__ small, int* __restrict__ small_count, int serial_max,
            int* __restrict__ first_ge_next = nullptr) {
  // (first_ge_next: the fused flag / scan kernel finds a node's first ">=" position with atomicMin and needs the
  //  word preset to the range's last position; whoever writes a node's next state presets it)
  const int node = nid[p];
  if (node < 0) { X2[p] = X[p]; Y2[p] = Y[p]; ID2[p] = ID[p]; return; }
  const KdfNode nd = nodes[node];
  const int lo = nd.a_lo, hi = nd.a_hi;
  const int k = node + ((nd.n_hi - node + 2) >> 1) - 1;  // ind_median = (num_points + 1) / 2   (:54)
  bool fin;
  int nlo = lo, nhi = hi;
  if (lo == hi) {                                        // :112
    fin = true;
    X2[p] = X[p]; Y2[p] = Y[p]; ID2[p] = ID[p];
  } else {
    const int base = F[lo];
    const int store = lo + (F[hi] - base);               // :127-138
    if (k < store) { fin = false; nhi = store - 1; }     // :149-153
    else if (k > store) { fin = false; nlo = store + 1; }
    else fin = true;
    if (!fin && nlo == nhi) fin = true;                  // the next call would return at :112 (k == nlo)
    if (p < lo || p > hi) {
      X2[p] = X[p]; Y2[p] = Y[p]; ID2[p] = ID[p];
    } else {
      const int pidx = (lo + hi) >> 1;
      if (p < hi && flag[p]) {                           // a "<" element: to the front, in encounter order
        const int src = kdf_virt(p, pidx, hi), dst = lo + (F[p] - base);
        X2[dst] = X[src]; Y2[dst] = Y[src]; ID2[dst] = ID[src];
      }
      if (p >= store) {                                  // destinations of the pivot and of the ">=" queue
        int src;
        if (p == store) src = pidx;                      // the pivot (:141-146)
        else {
          const int i0 = first_ge[node], f0 = F[i0];
          int e = (p == hi) ? store : p;                 // the queue's front goes to the end (:141-146)
          while (flag[e]) e = i0 + (F[e] - f0);
          src = kdf_virt(e, pidx, hi);
        }
        X2[p] = X[src]; Y2[p] = Y[src]; ID2[p] = ID[src];
      }
    }
  }
  // a child of at most kSerialMax points leaves the rounds: kdf_serial_kernel finishes its whole sub-tree
  const bool left_small = k - node <= serial_max, right_small = nd.n_hi - k <= serial_max;   // (0: none)
  if (p == hi) {  // one thread per node writes the next state(s)
    if (!fin) {
      nodes_next[node] = KdfNode{nlo, nhi, nd.n_hi, nd.depth}; *alive = 1;
      if (first_ge_next) first_ge_next[node] = nhi;
    } else {
      if (k > node) {                                                                                        // :64-65
        if (left_small) small[atomicAdd(small_count, 1)] = KdfSmall{node, k - 1, nd.depth + 1};
        else {
          nodes_next[node] = KdfNode{node, k - 1, k - 1, nd.depth + 1}; *alive = 1;
          if (first_ge_next) first_ge_next[node] = k - 1;
        }
      }
      if (k < nd.n_hi) {                                                                                     // :66-68
        if (right_small) small[atomicAdd(small_count, 1)] = KdfSmall{k + 1, nd.n_hi, nd.depth + 1};
        else {
          nodes_next[k + 1] = KdfNode{k + 1, nd.n_hi, nd.n_hi, nd.depth + 1}; *alive = 1;
          if (first_ge_next) first_ge_next[k + 1] = nd.n_hi;
        }
      }
    }
  }
  if (fin) {
    if (p == k) nid[p] = -1;      // :59-61 the node
    else if (p > k) nid[p] = right_small ? -2 : k + 1;
    else if (left_small) nid[p] = -2;
  }
}
template <class R>
__global__ void __launch_bounds__(256)
kdf_pass_kernel(const R* __restrict__ X, const R* __restrict__ Y, const int* __restrict__ ID, int* __restrict__ nid,
                const KdfNode* __restrict__ nodes, const int* __restrict__ flag, const int* __restrict__ F,
                const int* __restrict__ first_ge, int n, R* __restrict__ X2, R* __restrict__ Y2,
                int* __restrict__ ID2, KdfNode* __restrict__ nodes_next, int* __restrict__ alive,
                KdfSmall* __restrict__ small, int* __restrict__ small_count, int serial_max,
                int* __restrict__ first_ge_next) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  kdf_pass_at<R>(p, X, Y, ID, nid, nodes, flag, F, first_ge, X2, Y2, ID2, nodes_next, alive, small, small_count,
                 serial_max, first_ge_next);
}

// Flags, their prefix sums and every node's first ">=" position in ONE sweep (instead of flag kernel + three scan
// kernels + binary searches): a single-pass scan with decoupled look-back.  A tile (256 threads x 8 elements) takes
// its number from a ticket counter, publishes its aggregate, then walks back over its predecessors' words until one
// holds an inclusive prefix.  A word is stamped with the round it belongs to, so nothing is reset between rounds.
constexpr int kFusedItems = 8;
constexpr int kFusedTile = 256 * kFusedItems;
template <class R>
__global__ void __launch_bounds__(256)
kdf_flag_scan_kernel(const R* __restrict__ X, const R* __restrict__ Y, const int* __restrict__ nid,
                     const KdfNode* __restrict__ nodes, int n, int* __restrict__ flag, int* __restrict__ F,
                     int* __restrict__ first_ge, unsigned long long* state, unsigned* ticket, unsigned round,
                     int num_tiles) {
  __shared__ int s_tile, s_prefix;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = (int)(atomicAdd(ticket, 1u) - round * (unsigned)num_tiles);
  __syncthreads();
  const int tile = s_tile;
  // flags are computed with coalesced accesses (element j * 256 + tid of the tile) and handed through shared memory to
  // the thread that scans them (elements tid * 8 .. tid * 8 + 7)
  __shared__ int s_f[kFusedTile];
  const int tile0 = tile * kFusedTile;
#pragma unroll
  for (int j = 0; j < kFusedItems; ++j) {
    const int p = tile0 + j * 256 + tid;
    int fl = 0;
    if (p < n) {
      const int node = nid[p];
      if (node >= 0) {
        const KdfNode nd = nodes[node];
        if (p >= nd.a_lo && p < nd.a_hi) {
          const R* K = (nd.depth & 1) ? Y : X;                 // axis = mod(depth, 2)   (kd_tree.f90:45)
          const int pidx = (nd.a_lo + nd.a_hi) >> 1;           // :115
          fl = K[p == pidx ? nd.a_hi : p] < K[pidx] ? 1 : 0;   // :129 (strict)
          // i0 = first position of the pass whose element is not "<" (preset to a_hi by the previous pass)
          if (!fl && p < *(volatile int*)(first_ge + node)) atomicMin(first_ge + node, p);
        }
      }
      flag[p] = fl;
    }
    s_f[j * 256 + tid] = fl;
  }
  __syncthreads();
  int f[kFusedItems];
  int acc = 0;
#pragma unroll
  for (int i = 0; i < kFusedItems; ++i) { f[i] = s_f[tid * kFusedItems + i]; acc += f[i]; }
  int total;
  int run = block_exclusive_scan_256(acc, &total);
  if (tid == 0) {
    const unsigned long long stamp = (unsigned long long)(round & 0x3fffffffu) << 34;
    volatile unsigned long long* vs = state;
    if (tile == 0) {
      vs[0] = stamp | (2ull << 32) | (unsigned)total;
      s_prefix = 0;
    } else {
      vs[tile] = stamp | (1ull << 32) | (unsigned)total;       // my aggregate
      int excl = 0;
      for (int t = tile - 1;;) {
        const unsigned long long v = vs[t];
        const unsigned st = (unsigned)(v >> 32) & 3u;
        if ((v >> 34) != (stamp >> 34) || st == 0u) continue;  // not published in this round yet
        excl += (int)(unsigned)v;
        if (st == 2u) break;
        --t;
      }
      vs[tile] = stamp | (2ull << 32) | (unsigned)(excl + total);
      s_prefix = excl;
    }
  }
  __syncthreads();
  run += s_prefix;
#pragma unroll
  for (int i = 0; i < kFusedItems; ++i) { s_f[tid * kFusedItems + i] = run; run += f[i]; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kFusedItems; ++j) {
    const int p = tile0 + j * 256 + tid;
    if (p < n) F[p] = s_f[j * 256 + tid];
  }
}

// All rounds in ONE cooperative launch (arrays of up to one element per resident thread, ~150 000 points): a round of
// the host-driven form is eight sweeps over the array (flag, three scan kernels, first ">=", pass, two memsets) of
// ~10 us each however small the array.  Here every CTA owns a contiguous chunk of the array and the phases of a round are separated by grid-wide
// barriers: flags + chunk-local prefix | chunk offsets | first ">=" per node | pass.  Same device functions, same
// arithmetic, same result.  `alive` has three slots: round r reports in slot r % 3 and clears slot (r + 1) % 3, which
// nobody reads before the next barrier.
template <class R> struct KdfRounds {
  R *X, *Y, *X2, *Y2;
  int *ID, *ID2, *nid, *flag, *F, *first_ge, *totals, *alive, *small_count, *rounds;
  KdfNode *nodes, *nodes_next;
  KdfSmall* small;
  int n, serial_max, max_rounds;
};
template <class R>
__global__ void __launch_bounds__(256) kdf_rounds_kernel(KdfRounds<R> a) {
  cg::grid_group grid = cg::this_grid();
  const int nb = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  const int chunk = (a.n + nb - 1) / nb, per = (chunk + 255) >> 8;
  const int c0 = min(a.n, b * chunk), c1 = min(a.n, c0 + chunk);
  const int t0 = min(c1, c0 + tid * per), t1 = min(c1, t0 + per);   // this thread's elements (contiguous)
  R *X = a.X, *Y = a.Y, *X2 = a.X2, *Y2 = a.Y2;
  int *ID = a.ID, *ID2 = a.ID2;
  KdfNode *nodes = a.nodes, *nodes_next = a.nodes_next;
  int round = 0;
  for (;; ++round) {
    int* alive = a.alive + round % 3;
    if (b == 0 && tid == 0) a.alive[(round + 1) % 3] = 0;
    int acc = 0;
    for (int p = t0; p < t1; ++p) {
      const int f = kdf_flag_at<R>(p, X, Y, a.nid, nodes);
      a.flag[p] = f;
      acc += f;
    }
    int total;
    int run = block_exclusive_scan_256(acc, &total);
    for (int p = t0; p < t1; ++p) { a.F[p] = run; run += a.flag[p]; }
    if (tid == 0) a.totals[b] = total;
    grid.sync();
    int part = 0;
    for (int i = tid; i < b; i += 256) part += a.totals[i];
    int offset;
    block_exclusive_scan_256(part, &offset);
    if (offset)
      for (int p = t0; p < t1; ++p) a.F[p] += offset;
    grid.sync();
    for (int p = t0; p < t1; ++p) kdf_first_ge_at(p, a.nid, nodes, a.F, a.first_ge);
    grid.sync();
    for (int p = t0; p < t1; ++p)
      kdf_pass_at<R>(p, X, Y, ID, a.nid, nodes, a.flag, a.F, a.first_ge, X2, Y2, ID2, nodes_next, alive, a.small,
                     a.small_count, a.serial_max);
    grid.sync();
    { R* t = X; X = X2; X2 = t; t = Y; Y = Y2; Y2 = t; }
    { int* t = ID; ID = ID2; ID2 = t; }
    { KdfNode* t = nodes; nodes = nodes_next; nodes_next = t; }
    if (!*(volatile int*)alive || round + 1 >= a.max_rounds) break;
  }
  if (X != a.X)   // an odd number of rounds: the result goes back to the caller's first set of arrays
    for (int p = t0; p < t1; ++p) { a.X[p] = X[p]; a.Y[p] = Y[p]; a.ID[p] = ID[p]; }
  if (b == 0 && tid == 0) *a.rounds = round + 1;
}

// 1 iff two neighbours of a sorted key list are equal (-0.0 and +0.0 count as equal, as `<` sees them)
__global__ void __launch_bounds__(256)
kd_tie_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ tie) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 >= n) return;
  unsigned long long a = keys[i], b = keys[i + 1];
  if (a == 0x7fffffffffffffffull) a = 0x8000000000000000ull;
  if (b == 0x7fffffffffffffffull) b = 0x8000000000000000ull;
  if (a == b) *tie = 1;
}

// exclusive scan with caller-provided tile sums (the rounds below scan a few hundred times)
inline int kdf_scan(const int* in, int* out, int n, int* sums, cudaStream_t st) {
  const int tiles = div_up(n, kScanTile);
  if (tiles == 1) scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
  else if (tiles <= kScanSmallTiles) scan_small_kernel<<<1, kScanThreads, 0, st>>>(in, out, n);
  else {
    scan_reduce_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, sums);
    if (tiles <= kScanTile) scan_tile_kernel<<<1, kScanThreads, 0, st>>>(sums, sums, tiles, nullptr);
    else scan_small_kernel<<<1, kScanThreads, 0, st>>>(sums, sums, tiles);
    scan_tile_kernel<<<tiles, kScanThreads, 0, st>>>(in, out, n, sums);
  }
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

template <class R>
int kd_build_faithful(Scratch& scr, PointsView<R> pv, int n, int** order_dev) {
  cudaStream_t st = scr.stream();
  GI_REQUIRE((int64_t)n <= (int64_t)kScanTile * kScanTile * kScanSmallTiles, "too many points");
  R *X, *Y, *X2, *Y2;
  int *ID, *ID2, *nid, *flag, *F, *first_ge, *sums, *alive;
  KdfNode *nodes, *nodes_next;
  GI_TRY(scr.alloc(&X, (size_t)n)); GI_TRY(scr.alloc(&Y, (size_t)n)); GI_TRY(scr.alloc(&ID, (size_t)n));
  GI_TRY(scr.alloc(&X2, (size_t)n)); GI_TRY(scr.alloc(&Y2, (size_t)n)); GI_TRY(scr.alloc(&ID2, (size_t)n));
  GI_TRY(scr.alloc(&nid, (size_t)n)); GI_TRY(scr.alloc(&flag, (size_t)n)); GI_TRY(scr.alloc(&F, (size_t)n));
  GI_TRY(scr.alloc(&first_ge, (size_t)n));
  GI_TRY(scr.alloc(&nodes, (size_t)n)); GI_TRY(scr.alloc(&nodes_next, (size_t)n));
  GI_TRY(scr.alloc(&sums, (size_t)div_up(n, kScanTile)));
  GI_TRY(scr.alloc(&alive, 1));
  KdfSmall* small;
  int* small_count;
  GI_TRY(scr.alloc(&small, (size_t)n));
  GI_TRY(scr.alloc(&small_count, 1));
  GI_CUDA(cudaMemsetAsync(small_count, 0, sizeof(int), st));
  const int blocks = div_up(n, 256);
  kdf_init_kernel<R><<<blocks, 256, 0, st>>>(pv, n, X, Y, ID, nid, nodes, first_ge);
  static const bool no_serial = getenv("GI_KD_NO_SERIAL_FINISH") != nullptr;   // A/B and tests: rounds down to the leaves
  constexpr int kBatch = 8;  // rounds between two looks at the `alive` word
  if (n <= kSerialMax && !no_serial) {   // the whole tree is one small sub-array
    const KdfSmall whole{0, n - 1, 0};
    const int one = 1;
    GI_CUDA(cudaMemcpyAsync(small, &whole, sizeof(whole), cudaMemcpyHostToDevice, st));
    GI_CUDA(cudaMemcpyAsync(small_count, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    GI_CUDA(cudaStreamSynchronize(st));   // (the two host words go out of scope)
  }
  // all rounds in one cooperative launch where the device allows it (every CTA resident at once)
  static const bool no_coop = getenv("GI_KD_NO_COOP") != nullptr;   // A/B and tests: the host-driven rounds
  bool rounds_done = false;
  if ((n > kSerialMax || no_serial) && !no_coop) {
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    GI_CUDA(cudaGetDevice(&dev));
    GI_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    GI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    GI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kdf_rounds_kernel<R>, 256, 0));
    // (one element per resident thread: with several, every thread's dependent loads -- node record, prefix sums, hop
    //  chain -- queue up behind each other and a round of 1 M points takes 360 us instead of the launches' 110 us;
    //  measured: 10 000 points 8.5 -> 4.2 ms, 100 000 points 28 -> 16.5 ms, 1 M points 27 -> 86 ms)
    if (coop && per_sm > 0 && (int64_t)n <= (int64_t)sms * std::min(per_sm, 4) * 256) {
      const int grid = std::max(1, std::min(sms * std::min(per_sm, 4), div_up(n, 256)));
      int *totals, *alive3, *rounds;
      GI_TRY(scr.alloc(&totals, (size_t)grid));
      GI_TRY(scr.alloc(&alive3, 3));
      GI_TRY(scr.alloc(&rounds, 1));
      GI_CUDA(cudaMemsetAsync(alive3, 0, 3 * sizeof(int), st));
      KdfRounds<R> ra;
      ra.X = X; ra.Y = Y; ra.X2 = X2; ra.Y2 = Y2; ra.ID = ID; ra.ID2 = ID2; ra.nid = nid; ra.flag = flag; ra.F = F;
      ra.first_ge = first_ge; ra.totals = totals; ra.alive = alive3; ra.small_count = small_count; ra.rounds = rounds;
      ra.nodes = nodes; ra.nodes_next = nodes_next; ra.small = small;
      ra.n = n; ra.serial_max = no_serial ? 0 : kSerialMax;
      ra.max_rounds = (int)std::min<int64_t>(4 * (int64_t)n + 64, 0x7fffffff);
      void* kargs[] = {&ra};
      GI_CUDA(cudaLaunchCooperativeKernel((const void*)kdf_rounds_kernel<R>, dim3(grid), dim3(256), kargs, 0, st));
      rounds_done = true;
      if (getenv("GI_KD_TRACE")) {
        int h_rounds = 0;
        GI_CUDA(cudaMemcpyAsync(&h_rounds, rounds, sizeof(int), cudaMemcpyDeviceToHost, st));
        GI_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "[gridinterp] quickselect-faithful k-d build: n = %d, %d rounds (one cooperative launch, %d CTAs)\n",
                n, h_rounds, grid);
      }
    }
  }
  // host-driven rounds: two launches per round (flags + prefix sums + first ">=" in one sweep, then the pass)
  static const bool no_fused = getenv("GI_KD_NO_FUSED_SCAN") != nullptr;   // A/B and tests: six launches per round
  const int fused_tiles = div_up(n, kFusedTile);
  int* first_ge_next = nullptr;
  unsigned long long* tile_state = nullptr;
  unsigned* ticket = nullptr;
  if (!rounds_done && !no_fused) {
    GI_TRY(scr.alloc(&first_ge_next, (size_t)n));
    GI_TRY(scr.alloc(&tile_state, (size_t)fused_tiles));
    GI_TRY(scr.alloc(&ticket, 1));
    GI_CUDA(cudaMemsetAsync(tile_state, 0xff, sizeof(unsigned long long) * (size_t)fused_tiles, st));
    GI_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  }
  for (int64_t round = 0; (n > kSerialMax || no_serial) && !rounds_done; round += kBatch) {
    GI_REQUIRE(round <= 4 * (int64_t)n + 64, "k-d build did not converge");
    GI_CUDA(cudaMemsetAsync(alive, 0, sizeof(int), st));
    for (int b = 0; b < kBatch; ++b) {
      if (no_fused) {
        kdf_flag_kernel<R><<<blocks, 256, 0, st>>>(X, Y, nid, nodes, n, flag);
        GI_TRY(kdf_scan(flag, F, n, sums, st));
        kdf_first_ge_kernel<<<blocks, 256, 0, st>>>(nid, nodes, F, n, first_ge);
      } else {
        kdf_flag_scan_kernel<R><<<fused_tiles, 256, 0, st>>>(X, Y, nid, nodes, n, flag, F, first_ge, tile_state, ticket,
                                                            (unsigned)(round + b), fused_tiles);
      }
      if (b == kBatch - 1) GI_CUDA(cudaMemsetAsync(alive, 0, sizeof(int), st));  // what the LAST round leaves
      kdf_pass_kernel<R><<<blocks, 256, 0, st>>>(X, Y, ID, nid, nodes, flag, F, first_ge, n, X2, Y2, ID2,
                                                 nodes_next, alive, small, small_count, no_serial ? 0 : kSerialMax,
                                                 no_fused ? nullptr : first_ge_next);
      std::swap(X, X2); std::swap(Y, Y2); std::swap(ID, ID2); std::swap(nodes, nodes_next);
      if (!no_fused) std::swap(first_ge, first_ge_next);
    }
    int h_alive = 0;
    GI_CUDA(cudaMemcpyAsync(&h_alive, alive, sizeof(int), cudaMemcpyDeviceToHost, st));
    GI_CUDA(cudaStreamSynchronize(st));
    if (!h_alive) {
      if (getenv("GI_KD_TRACE")) fprintf(stderr, "[gridinterp] quickselect-faithful k-d build: n = %d, %lld rounds\n", n, (long long)(round + kBatch));
      break;
    }
  }
  kdf_serial_kernel<R><<<kNumSMs * 8, 32, 0, st>>>(X, Y, ID, small, small_count);
  GI_CUDA(cudaGetLastError());
  *order_dev = ID;
  return GI_OK;
}

// builds the implicit tree; *order_dev (n ints, 0-based point ids in tree order) lives in scr
template <class R>
int kd_build(Scratch& scr, PointsView<R> pv, int n, int flags, int** order_dev) {
  cudaStream_t st = scr.stream();
  if (flags & GI_KD_FAITHFUL_BUILD) return kd_build_faithful<R>(scr, pv, n, order_dev);
  unsigned long long *keys, *keys_tmp;
  int* tie;
  GI_TRY(scr.alloc(&tie, 1));
  GI_CUDA(cudaMemsetAsync(tie, 0, sizeof(int), st));
  int *list[2], *pos[2], *list_new, *pos_new, *seg_lo, *seg_hi, *seg_lo_new, *seg_hi_new, *flag, *ids_tmp, *hist;
  GI_TRY(scr.alloc(&keys, (size_t)n));
  GI_TRY(scr.alloc(&keys_tmp, (size_t)n));
  GI_TRY(scr.alloc(&ids_tmp, (size_t)n));
  GI_TRY(scr.alloc(&hist, radix_hist_size<unsigned long long>(n)));
  GI_TRY(scr.alloc(&list[0], (size_t)n));
  GI_TRY(scr.alloc(&list[1], (size_t)n));
  GI_TRY(scr.alloc(&pos[0], (size_t)n));
  GI_TRY(scr.alloc(&pos[1], (size_t)n));
  GI_TRY(scr.alloc(&list_new, (size_t)n));
  GI_TRY(scr.alloc(&pos_new, (size_t)n));
  GI_TRY(scr.alloc(&seg_lo, (size_t)n));
  GI_TRY(scr.alloc(&seg_hi, (size_t)n));
  GI_TRY(scr.alloc(&seg_lo_new, (size_t)n));
  GI_TRY(scr.alloc(&seg_hi_new, (size_t)n));
  GI_TRY(scr.alloc(&flag, (size_t)n));
  for (int axis = 0; axis < 2; ++axis) {
    sort_keys_kernel<R><<<div_up(n, 256), 256, 0, st>>>(pv, n, axis, keys, list[axis]);
    unsigned long long* k_sorted; int* l_sorted;  // 8 passes: back in keys / list[axis]
    GI_TRY(radix_sort<unsigned long long>(keys, list[axis], keys_tmp, ids_tmp, hist, n, 8, st, &k_sorted, &l_sorted));
    kd_tie_kernel<<<div_up(n, 256), 256, 0, st>>>(k_sorted, n, tie);
  }
  // tied coordinates: the reference's tree depends on its quickselect's swap history (see kd_build_faithful)
  int h_tie = 0;
  GI_CUDA(cudaMemcpyAsync(&h_tie, tie, sizeof(int), cudaMemcpyDeviceToHost, st));
  GI_CUDA(cudaStreamSynchronize(st));
  if (h_tie) return kd_build_faithful<R>(scr, pv, n, order_dev);
  const int blocks = div_up(n, 256);
  kd_init_kernel<<<blocks, 256, 0, st>>>(list[0], list[1], n, pos[0], pos[1], seg_lo, seg_hi);
  int levels = 0;
  while ((1ll << levels) <= (long long)n) ++levels;  // floor(log2 n) + 1
  int d0 = 0;   // levels done by launches: until every segment fits a CTA's shared memory (kd_finish_kernel)
  while ((n >> d0) > kFinishMax) ++d0;
  static const bool no_finish = getenv("GI_KD_NO_SMEM_FINISH") != nullptr;   // A/B and tests
  if (no_finish) d0 = levels;
  for (int d = 0; d < d0; ++d) {
    const int ax = d & 1, ot = ax ^ 1;  // axis = mod(depth, 2)   (kd_tree.f90:45)
    kd_flag_kernel<<<blocks, 256, 0, st>>>(list[ot], pos[ax], seg_lo, seg_hi, n, flag);
    GI_TRY(exclusive_scan_i32(scr, flag, flag, n));
    kd_partition_kernel<<<blocks, 256, 0, st>>>(list[ax], list[ot], pos[ax], pos[ot], seg_lo, seg_hi, flag, n,
                                                list_new, pos_new, seg_lo_new, seg_hi_new);
    std::swap(list[ot], list_new);
    std::swap(pos[ot], pos_new);
    std::swap(seg_lo, seg_lo_new);
    std::swap(seg_hi, seg_hi_new);
  }
  if (d0 < levels) kd_finish_kernel<<<1u << d0, 256, 0, st>>>(list[0], list[1], pos[0], pos[1], n, d0, levels);
  GI_CUDA(cudaGetLastError());
  *order_dev = list[0];  // both lists are identical once every position is a node
  return GI_OK;
}

// ---------------------------------------------------------------------------------------------
// query + IDW
// ---------------------------------------------------------------------------------------------
template <class R> struct KdArgs {
  TargetWindow<R> win;
  const XY<R>* nodes;  // coordinates in implicit-tree order
  const R* nval;       // values in tree order
  const int* order;    // 0-based point id per tree position
  int n, k;
  R* out;
  int32_t* nn_index;   // optional (window, k), 1-based ids, 0 = slot never filled
  R* nn_dist2;         // optional
  int64_t* counters;   // [0] node visits
  int* status;
  R* tk_d;             // k > 32: neighbour lists in scratch memory (topk.cuh, KC = 0), k x tk_stride
  int* tk_id;
  int64_t tk_stride;
  PointsView<R> tgt;   // scattered targets (num_tgt > 0): thread i takes target i instead of a grid node
  int64_t num_tgt;
};

template <class R> struct alignas(16) StackEntry {
  int lo, hi;
  R key;  // what the prune rule compares with max d2 (see prune_key); sign bit = parity of the subtree's depth
};

// PRUNE: 0 = reference rule + state-preserving skip (default), 1 = reference rule only (visit-count parity),
// 2 = exact rule diff*diff < max d2 (extension: exact kNN).
// The far branch behind a split at distance ad = |diff| is visited iff prune_key(ad) < max d2:
//   1: ad < max d2                              (kd_tree.f90:191,198)
//   2: ad*ad < max d2
//   0: ad < max d2 AND ad*ad < max d2  <=>  max(ad, ad*ad) < max d2, and max(ad, ad*ad) is ad*ad iff ad >= 1
// sign-bit helpers on the integer pipe (fabs / negation of a double are FP64-pipe DADDs otherwise)
__device__ __forceinline__ double clear_sign(double x) {
  return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}
__device__ __forceinline__ float clear_sign(float x) { return __int_as_float(__float_as_int(x) & 0x7fffffff); }
__device__ __forceinline__ double or_sign(double x, unsigned sign) {
  return __hiloint2double(__double2hiint(x) | (int)sign, __double2loint(x));
}
__device__ __forceinline__ float or_sign(float x, unsigned sign) { return __int_as_float(__float_as_int(x) | (int)sign); }
__device__ __forceinline__ unsigned sign_of(double x) { return (unsigned)__double2hiint(x) & 0x80000000u; }
__device__ __forceinline__ unsigned sign_of(float x) { return (unsigned)__float_as_int(x) & 0x80000000u; }
__device__ __forceinline__ bool at_least_one(double ad) { return __double2hiint(ad) >= 0x3ff00000; }  // ad >= 0
__device__ __forceinline__ bool at_least_one(float ad) { return ad >= 1.0f; }

template <class R, int PRUNE> __device__ __forceinline__ R prune_key(R ad) {
  if (PRUNE == 1) return ad;
  if (PRUNE == 2) return ad * ad;
  return at_least_one(ad) ? ad * ad : ad;
}

// ncu on the first version (profiles/r01_kd_query_ncu.csv): 513 warp instructions per target at 12 of 32 lanes
// active -- a "descend to a leaf, then pop" loop nest makes every lane wait for the longest descent of its warp
// in every round, and the position-counting insertion cost ~110 instructions.  This version is one flat loop
// whose body is "pop if the current segment is empty; visit one node": lanes advance independently and meet
// again at every node visit.  Far branches that the prune rule already rejects when they are created are never
// pushed (max d2 only shrinks, so the rejection is final), which removes most stack traffic.
template <class R, int KC, int PRUNE>
__global__ void __launch_bounds__(128) kd_query_kernel(const KdArgs<R> a) {
  const TargetWindow<R>& w = a.win;
  // a warp owns a compact 8 x 4 patch of targets (8 along the contiguous output axis): neighbouring targets
  // descend through the same nodes, which keeps the 32 traversals as coherent as a per-thread walk can be
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int f = w.f0 + blockIdx.x * 32 + wid * 8 + (lane & 7);
  const int s = w.s0 + blockIdx.y * 4 + (lane >> 3);
  const bool scattered = a.num_tgt > 0;  // the reference is grid-only (interpolation.f90:154-158); SURVEY 8f
  const int64_t ti = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = scattered ? ti < a.num_tgt : (f < w.f1 && s < w.s1);
  unsigned visits = 0;
  if (active) {
    R tx, ty;
    if (scattered) {
      a.tgt.xy(ti, tx, ty);
    } else {
      const int ix = w.fast_is_y ? s : f, iy = w.fast_is_y ? f : s;
      tx = __ldg((w.fast_is_y ? w.slow_axis : w.fast_axis) + ix);         // interpolation.f90:158
      ty = __ldg((w.fast_is_y ? w.fast_axis : w.slow_axis) + iy);
    }
    TopK<R, KC> tk;
    if constexpr (KC == 0)
      tk.bind(a.tk_d, a.tk_id, a.tk_stride, ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x);
    tk.init(a.k, Consts<R>::sentinel(), -1);                              // :159
    // Deferred far branches, at most one per tree level.  Measured alternatives (profiles/r01_kd_esrg_variants.md):
    // top entry kept in registers +16 %, 48 registers for 10 CTAs/SM +8 % (spills), two entries per back-up
    // round -1 %.
    StackEntry<R> stack[32];
    int sp = 0;
    int lo = 0, hi = a.n;
    unsigned odd = 0u;  // 0x80000000 at odd depth, else 0: axis = mod(depth, 2)   (kd_tree.f90:45)
    for (;;) {
      if (lo >= hi) {                                                     // :174 -- back up to a deferred branch
        bool got = false;
        while (sp > 0) {
          const StackEntry<R> e = stack[--sp];
          if (clear_sign(e.key) < tk.dlast) {                             // :191,:198 (see prune_key)
            lo = e.lo; hi = e.hi;
            odd = sign_of(e.key);
            got = true;
            break;
          }
        }
        if (!got) break;
      }
      const int m = lo + ((hi - lo + 1) >> 1) - 1;
      const XY<R> nd = a.nodes[m];
      const R ddx = nd.x - tx, ddy = nd.y - ty;
      const R d2 = ddx * ddx + ddy * ddy;                                 // :177,:249
      if (d2 < tk.dlast) tk.template insert<false>(d2, m);                // :222-236
      ++visits;
      // diff = t(axis) - node(axis) (:183-184) is exactly -(node - t): reuse the difference computed for d2
      const R nmt = odd ? ddy : ddx;
      int flo, fhi;
      if (nmt > R(0)) { flo = m + 1; fhi = hi; hi = m; }                  // :188 diff < 0: near = left
      else { flo = lo; fhi = m; lo = m + 1; }                             // :195 near = right
      odd ^= 0x80000000u;
      const R key = prune_key<R, PRUNE>(clear_sign(nmt));
      if (flo < fhi && key < tk.dlast && sp < 32) {
        StackEntry<R> e;
        e.lo = flo; e.hi = fhi;
        e.key = or_sign(key, odd);
        stack[sp++] = e;
      }
    }
    // IDW (interpolation.f90:164-178); rank i of the list lives in slot tk.base + i
    const int64_t o = scattered ? ti : (int64_t)(s - w.s0) * w.nf() + (f - w.f0);
    R numerator = R(0), denominator = R(0), first_val = R(0), first_d2 = R(0);
    auto take = [&](int rank, int pos, R d2r) {
      const R dist = sqrt(d2r);
      const R v = pos >= 0 ? __ldg(a.nval + pos) : R(0);  // unfilled slot: uninitialised read in the reference
      if (rank == 0) { first_val = v; first_d2 = d2r; }
      numerator = numerator + (v / dist);
      denominator = denominator + (R(1) / dist);
      if (a.nn_index) a.nn_index[o * a.k + rank] = pos >= 0 ? __ldg(a.order + pos) + 1 : 0;
      if (a.nn_dist2) a.nn_dist2[o * a.k + rank] = d2r;
    };
    if constexpr (KC == 0) {
      for (int j = 0; j < a.k; ++j) take(j, tk.index(j), tk.dist2(j));
    } else {
#pragma unroll
      for (int j = 0; j < KC; ++j)
        if (j >= tk.base) take(j - tk.base, tk.id[j], tk.d[j]);
    }
    a.out[o] = (sqrt(first_d2) <= Consts<R>::tiny_e4()) ? first_val : numerator / denominator;
  }
  if (a.counters) {
    unsigned long long v = visits;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long*)a.counters, v);
  }
}

template <class R, int KC>
void launch_query(const KdArgs<R>& a, int prune, dim3 grid, cudaStream_t st) {
  if (prune == 2) kd_query_kernel<R, KC, 2><<<grid, 128, 0, st>>>(a);
  else if (prune == 1) kd_query_kernel<R, KC, 1><<<grid, 128, 0, st>>>(a);
  else kd_query_kernel<R, KC, 0><<<grid, 128, 0, st>>>(a);
}

// build + gather: the implicit tree as the query reads it (lives in `scr`)
template <class R> struct KdTreeDev {
  XY<R>* nodes;
  R* nval;
  int* order;
  int n;
};
template <class R>
int kd_prepare(Scratch& keep, Scratch& tmp, const R* points, int num_points, const R* data_in, int flags,
               KdTreeDev<R>* t) {
  cudaStream_t st = tmp.stream();
  const PointsView<R> pv = points_2n(points, num_points, flags & GI_POINTS_ROWMAJOR);
  int* order_tmp;
  GI_TRY(kd_build<R>(tmp, pv, num_points, flags, &order_tmp));
  GI_TRY(keep.alloc(&t->order, (size_t)num_points));
  GI_CUDA(cudaMemcpyAsync(t->order, order_tmp, sizeof(int) * (size_t)num_points, cudaMemcpyDeviceToDevice, st));
  GI_TRY(keep.alloc(&t->nodes, (size_t)num_points));
  GI_TRY(keep.alloc(&t->nval, (size_t)num_points));
  t->n = num_points;
  kd_gather_kernel<R><<<div_up(num_points, 256), 256, 0, st>>>(pv, data_in, t->order, num_points, t->nodes,
                                                               data_in ? t->nval : nullptr, nullptr);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}
template <class R> int kd_prepare(Scratch& scr, const R* points, int n, const R* data_in, int flags, KdTreeDev<R>* t) {
  return kd_prepare<R>(scr, scr, points, n, data_in, flags, t);
}

// query + IDW for the targets rows [row0,row1) x columns [col0,col1); `out` (and nn_*) are window-shaped
template <class R>
int kd_query_window(const KdTreeDev<R>& t, const R* x_axis, int len_x, const R* y_axis, int len_y, int k, int row0,
                    int row1, int col0, int col1, R* out, int32_t* nn_index, R* nn_dist2, int64_t* counters,
                    int flags, int* status, cudaStream_t st) {
  if (row0 >= row1 || col0 >= col1) return GI_OK;
  KdArgs<R> a;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  a.win = rowmajor ? TargetWindow<R>{x_axis, y_axis, len_x, len_y, col0, col1, row0, row1, 0}
                   : TargetWindow<R>{y_axis, x_axis, len_y, len_x, row0, row1, col0, col1, 1};
  a.nodes = t.nodes; a.nval = t.nval; a.order = t.order; a.n = t.n; a.k = k;
  a.out = out; a.nn_index = nn_index; a.nn_dist2 = nn_dist2;
  a.counters = counters; a.status = status;
  dim3 grid(div_up(a.win.nf(), 32), div_up(a.win.ns(), 4));
  GI_REQUIRE(grid.y <= 65535, "window too tall for one launch");
  const int prune = (flags & GI_KD_EXACT_PRUNE) ? 2 : ((flags & GI_KD_REFERENCE_VISITS) ? 1 : 0);
  a.tk_d = nullptr; a.tk_id = nullptr; a.tk_stride = 0;
  a.tgt = PointsView<R>{nullptr, 0, 0}; a.num_tgt = 0;
  if (k <= 4) launch_query<R, 4>(a, prune, grid, st);
  else if (k <= 8) launch_query<R, 8>(a, prune, grid, st);
  else if (k <= 16) launch_query<R, 16>(a, prune, grid, st);
  else if (k <= 32) launch_query<R, 32>(a, prune, grid, st);
  else {
    // any k: the lists live in scratch memory (topk.cuh, KC = 0); the window is done in slices of slow rows so
    // that the scratch stays below ~1 GB
    Scratch scr(st);
    const int64_t per_row_block = (int64_t)grid.x * 128;  // threads of one grid row (4 slow rows)
    const int64_t max_threads = std::max<int64_t>(per_row_block, ((int64_t)1 << 30) / ((int64_t)k * (sizeof(R) + 4)));
    const int blocks_y = (int)std::min<int64_t>(grid.y, std::max<int64_t>(1, max_threads / per_row_block));
    a.tk_stride = per_row_block * blocks_y;
    GI_TRY(scr.alloc(&a.tk_d, (size_t)a.tk_stride * k));
    GI_TRY(scr.alloc(&a.tk_id, (size_t)a.tk_stride * k));
    const TargetWindow<R> full = a.win;
    R* const out0 = a.out; int32_t* const nn0 = a.nn_index; R* const nd0 = a.nn_dist2;
    for (int by = 0; by < (int)grid.y; by += blocks_y) {
      const int s_lo = full.s0 + by * 4, s_hi = std::min(full.s1, s_lo + blocks_y * 4);
      a.win.s0 = s_lo; a.win.s1 = s_hi;
      const int64_t off = (int64_t)(s_lo - full.s0) * full.nf();
      a.out = out0 + off;
      a.nn_index = nn0 ? nn0 + off * k : nullptr;
      a.nn_dist2 = nd0 ? nd0 + off * k : nullptr;
      launch_query<R, 0>(a, prune, dim3(grid.x, div_up(s_hi - s_lo, 4)), st);
    }
  }
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

// the same query for scattered targets (m, 2): thread i takes target i
template <class R>
int kd_query_points(const KdTreeDev<R>& t, const R* targets, int64_t m, int k, R* out, int32_t* nn_index, R* nn_dist2,
                    int flags, int* status, cudaStream_t st) {
  if (m <= 0) return GI_OK;
  KdArgs<R> a;
  a.win = TargetWindow<R>{nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0};
  a.nodes = t.nodes; a.nval = t.nval; a.order = t.order; a.n = t.n; a.k = k;
  a.counters = nullptr; a.status = status;
  a.tk_d = nullptr; a.tk_id = nullptr; a.tk_stride = 0;
  const int prune = (flags & GI_KD_EXACT_PRUNE) ? 2 : ((flags & GI_KD_REFERENCE_VISITS) ? 1 : 0);
  Scratch scr(st);
  // slices of at most 2^30 / (k * 12) targets when the lists live in scratch memory (k > 32), else 2^30 targets
  int64_t slice = (int64_t)1 << 30;
  if (k > 32) {
    slice = std::max<int64_t>(128, (((int64_t)1 << 30) / ((int64_t)k * (sizeof(R) + 4))) / 128 * 128);
    a.tk_stride = std::min<int64_t>(slice, (m + 127) / 128 * 128);
    GI_TRY(scr.alloc(&a.tk_d, (size_t)a.tk_stride * k));
    GI_TRY(scr.alloc(&a.tk_id, (size_t)a.tk_stride * k));
  }
  for (int64_t i0 = 0; i0 < m; i0 += slice) {
    const int64_t cnt = std::min(slice, m - i0);
    PointsView<R> tv = points_n2(targets, m, flags & GI_TARGETS_ROWMAJOR);
    tv.base += i0 * tv.stride;
    a.tgt = tv; a.num_tgt = cnt;
    a.out = out + i0;
    a.nn_index = nn_index ? nn_index + i0 * k : nullptr;
    a.nn_dist2 = nn_dist2 ? nn_dist2 + i0 * k : nullptr;
    const dim3 grid((unsigned)div_up(cnt, 128), 1);
    if (k <= 4) launch_query<R, 4>(a, prune, grid, st);
    else if (k <= 8) launch_query<R, 8>(a, prune, grid, st);
    else if (k <= 16) launch_query<R, 16>(a, prune, grid, st);
    else if (k <= 32) launch_query<R, 32>(a, prune, grid, st);
    else launch_query<R, 0>(a, prune, grid, st);
  }
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

}  // namespace

// IDW from the k nearest neighbours (interpolation.f90:107-194) at scattered targets (m, 2) instead of grid nodes
// (the reference is grid-only, :154-158): same tree, same traversal, same sums.
template <class R>
int kdtree_points_device(const R* points, int num_points, const R* data_in, const R* targets, int64_t num_targets, int k,
                         R* data_ip, int32_t* nn_index, R* nn_dist2, int flags, cudaStream_t st) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_REQUIRE(num_points >= 1 && num_targets >= 0, "bad dimensions");
  GI_TRY(ensure_device());
  if (num_targets == 0) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("kd_build");
  KdTreeDev<R> tree;
  GI_TRY(kd_prepare<R>(scr, points, num_points, data_in, flags, &tree));
  ph.mark("kd_query");
  GI_TRY(kd_query_points<R>(tree, targets, num_targets, k, data_ip, nn_index, nn_dist2, flags, slot.dev, st));
  ph.finish();
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

template <class R>
int kdtree_device(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x,
                  const R* y_axis, int len_y, int k, int row_begin, int row_end, R* data_ip,
                  int32_t* nn_index, R* nn_dist2, int64_t* counters_out, int flags, cudaStream_t st) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_TRY(ensure_device());
  if (row_begin == row_end) return GI_OK;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("kd_build");
  KdTreeDev<R> tree;
  GI_TRY(kd_prepare<R>(scr, points, num_points, data_in, flags, &tree));
  int64_t* counters = nullptr;
  if (counters_out) {
    GI_TRY(scr.alloc(&counters, 4));
    GI_CUDA(cudaMemsetAsync(counters, 0, sizeof(int64_t) * 4, st));
  }
  ph.mark("kd_query");
  GI_TRY(kd_query_window<R>(tree, x_axis, len_x, y_axis, len_y, k, row_begin, row_end, 0, len_x, data_ip, nn_index,
                            nn_dist2, counters, flags, slot.dev, st));
  ph.finish();
  if (counters_out)
    GI_CUDA(cudaMemcpyAsync(counters_out, counters, sizeof(int64_t) * 4, cudaMemcpyDeviceToDevice, st));
  GI_TRY(flush_status(st, slot));
  return GI_OK;
}

namespace {
// output bands computed on the compute stream and copied back on the copy stream (common.cuh); dout is the
// device twin of data_ip (host), both shaped (row_end-row_begin, len_x) in the layout `flags` selects
template <class R>
int kd_run_banded(HostPipe* pipe, const KdTreeDev<R>& tree, const R* dx, int len_x, const R* dy, int len_y, int k,
                  int row_begin, int row_end, R* dout, R* data_ip, int flags, const StatusSlot& slot) {
  cudaStream_t st = pipe->compute;
  const int nrows = row_end - row_begin;
  const bool rowmajor = flags & GI_FIELD_ROWMAJOR;
  // bands along the axis that is NOT contiguous in memory: each band is one contiguous piece of the output
  const int64_t lo = rowmajor ? row_begin : 0, hi = rowmajor ? row_end : len_x;
  const int64_t line = rowmajor ? len_x : nrows;  // elements per unit of the banded axis
  const int nb = plan_bands(hi - lo, (int64_t)nrows * len_x * sizeof(R));
  for (int b = 0; b < nb; ++b) {
    int64_t b0, b1;
    band_range(lo, hi, nb, b, 4, &b0, &b1);
    if (b0 >= b1) continue;
    R* band_out = dout + (b0 - lo) * line;
    if (rowmajor)
      GI_TRY(kd_query_window<R>(tree, dx, len_x, dy, len_y, k, (int)b0, (int)b1, 0, len_x, band_out, nullptr, nullptr,
                                nullptr, flags, slot.dev, st));
    else
      GI_TRY(kd_query_window<R>(tree, dx, len_x, dy, len_y, k, row_begin, row_end, (int)b0, (int)b1, band_out,
                                nullptr, nullptr, nullptr, flags, slot.dev, st));
    if (nb > 1) {  // (a single band needs no second stream: small calls are latency-bound)
      GI_CUDA(cudaEventRecord(pipe->computed[b], st));
      GI_CUDA(cudaStreamWaitEvent(pipe->copy_out, pipe->computed[b], 0));
    }
    GI_TRY(to_host(nb > 1 ? pipe->copy_out : st, data_ip + (b0 - lo) * line, band_out, (size_t)((b1 - b0) * line)));
  }
  GI_TRY(flush_status(st, slot));
  if (nb > 1) GI_CUDA(cudaStreamSynchronize(pipe->copy_out));
  return GI_OK;
}
}  // namespace

// GI_HOST form without debug outputs: inputs in, tree built once, output bands computed and copied back in a
// pipeline (common.cuh).  host pointers in, synchronous.
template <class R>
int kdtree_host(const R* points, int num_points, const R* data_in, const R* x_axis, int len_x, const R* y_axis,
                int len_y, int k, int row_begin, int row_end, R* data_ip, int flags) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_TRY(ensure_device());
  if (row_begin == row_end) return GI_OK;
  HostPipe* pipe;
  GI_TRY(host_pipe(&pipe));
  cudaStream_t st = pipe->compute;
  StatusSlot slot;
  GI_TRY(get_status_slot(st, &slot));
  Scratch scr(st);
  Phases ph(st);
  ph.mark("h2d");
  R *dp, *dd, *dx, *dy, *dout;
  GI_TRY(to_device(scr, points, (size_t)2 * num_points, &dp));
  GI_TRY(to_device(scr, data_in, (size_t)num_points, &dd));
  GI_TRY(to_device(scr, x_axis, (size_t)len_x, &dx));
  GI_TRY(to_device(scr, y_axis, (size_t)len_y, &dy));
  const int nrows = row_end - row_begin;
  GI_TRY(scr.alloc(&dout, (size_t)nrows * len_x));
  ph.mark("kd_build");
  KdTreeDev<R> tree;
  GI_TRY(kd_prepare<R>(scr, dp, num_points, dd, flags, &tree));
  ph.mark("kd_query");
  GI_TRY(kd_run_banded<R>(pipe, tree, dx, len_x, dy, len_y, k, row_begin, row_end, dout, data_ip, flags, slot));
  ph.finish();
  return read_status(st, true);
}

// ---------------------------------------------------------------------------------------------
// plan: tree + target grid resident on the device, one execute() per field of nodal values
// ---------------------------------------------------------------------------------------------
namespace {
template <class R> struct KdPlan : PlanBase {
  Scratch keep{nullptr, true};
  KdTreeDev<R> tree;
  R *dx, *dy;
  int len_x, len_y, k, row_begin, row_end, flags;

  int execute(const void* data_in_v, void* data_ip_v, int mem, cudaStream_t st_in) override {
    const R* data_in = static_cast<const R*>(data_in_v);
    R* data_ip = static_cast<R*>(data_ip_v);
    if (row_begin == row_end) return GI_OK;
    HostPipe* pipe = nullptr;
    cudaStream_t st = st_in;
    if (mem == GI_HOST) { GI_TRY(host_pipe(&pipe)); st = pipe->compute; }
    StatusSlot slot;
    GI_TRY(get_status_slot(st, &slot));
    Scratch scr(st);
    Phases ph(st);
    const R* dd = data_in;
    R* dout = data_ip;
    if (mem == GI_HOST) {
      ph.mark("h2d");
      R* tmp;
      GI_TRY(to_device(scr, data_in, (size_t)tree.n, &tmp));
      dd = tmp;
      GI_TRY(scr.alloc(&dout, (size_t)(row_end - row_begin) * len_x));
    }
    ph.mark("kd_values");
    kd_gather_kernel<R><<<div_up(tree.n, 256), 256, 0, st>>>(PointsView<R>{nullptr, 0, 0}, dd, tree.order, tree.n, nullptr,
                                                            tree.nval, nullptr);
    GI_CUDA(cudaGetLastError());
    ph.mark("kd_query");
    if (mem == GI_HOST) {
      GI_TRY(kd_run_banded<R>(pipe, tree, dx, len_x, dy, len_y, k, row_begin, row_end, dout, data_ip, flags, slot));
      ph.finish();
      return read_status(st, true);
    }
    GI_TRY(kd_query_window<R>(tree, dx, len_x, dy, len_y, k, row_begin, row_end, 0, len_x, dout, nullptr, nullptr,
                              nullptr, flags, slot.dev, st));
    ph.finish();
    return flush_status(st, slot);
  }
};
}  // namespace

template <class R>
int kdtree_plan_create(const R* points, int num_points, const R* x_axis, int len_x, const R* y_axis, int len_y, int k,
                       int row_begin, int row_end, int flags, int mem, cudaStream_t st_in, PlanBase** out) {
  GI_REQUIRE(k >= 1, "num_neighbours must be at least 1");
  GI_TRY(ensure_device());
  cudaStream_t st = st_in;
  if (mem == GI_HOST) { HostPipe* pipe; GI_TRY(host_pipe(&pipe)); st = pipe->compute; }
  std::unique_ptr<KdPlan<R>> plan(new KdPlan<R>());
  plan->real_bytes = sizeof(R);
  GI_CUDA(cudaGetDevice(&plan->device));
  plan->len_x = len_x; plan->len_y = len_y; plan->k = k; plan->row_begin = row_begin; plan->row_end = row_end;
  plan->flags = flags;
  Scratch tmp(st);
  const R* dp = points;
  GI_TRY(plan->keep.alloc(&plan->dx, (size_t)len_x));
  GI_TRY(plan->keep.alloc(&plan->dy, (size_t)len_y));
  const cudaMemcpyKind kind = mem == GI_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  GI_CUDA(cudaMemcpyAsync(plan->dx, x_axis, sizeof(R) * (size_t)len_x, kind, st));
  GI_CUDA(cudaMemcpyAsync(plan->dy, y_axis, sizeof(R) * (size_t)len_y, kind, st));
  if (mem == GI_HOST) {
    R* t;
    GI_TRY(to_device(tmp, points, (size_t)2 * num_points, &t));
    dp = t;
  }
  GI_TRY(kd_prepare<R>(plan->keep, tmp, dp, num_points, (const R*)nullptr, flags, &plan->tree));
  GI_CUDA(cudaStreamSynchronize(st));  // the plan may be used from any stream afterwards
  *out = plan.release();
  return GI_OK;
}

template <class R>
int kd_build_device(const R* points, int num_points, int32_t* order_out, int flags, cudaStream_t st) {
  GI_TRY(ensure_device());
  Scratch scr(st);
  const PointsView<R> pv = points_2n(points, num_points, flags & GI_POINTS_ROWMAJOR);
  int* order;
  GI_TRY(kd_build<R>(scr, pv, num_points, flags, &order));
  kd_gather_kernel<R><<<div_up(num_points, 256), 256, 0, st>>>(pv, nullptr, order, num_points, nullptr, nullptr,
                                                               order_out);
  GI_CUDA(cudaGetLastError());
  return GI_OK;
}

#define GI_INSTANTIATE(R)                                                                                       \
  template int kdtree_device<R>(const R*, int, const R*, const R*, int, const R*, int, int, int, int, R*, int32_t*, \
                                R*, int64_t*, int, cudaStream_t);                                               \
  template int kd_build_device<R>(const R*, int, int32_t*, int, cudaStream_t);                                  \
  template int kdtree_points_device<R>(const R*, int, const R*, const R*, int64_t, int, R*, int32_t*, R*, int,  \
                                       cudaStream_t);                                                           \
  template int kdtree_host<R>(const R*, int, const R*, const R*, int, const R*, int, int, int, int, R*, int);    \
  template int kdtree_plan_create<R>(const R*, int, const R*, int, const R*, int, int, int, int, int, int,       \
                                     cudaStream_t, PlanBase**);
GI_INSTANTIATE(double)
GI_INSTANTIATE(float)

}  // namespace gi
