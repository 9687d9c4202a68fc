// Residue k-NN graph construction on the device: the reference's
// protein_to_graph_unbound_bound_residuesonly (src/utils/protein_utils.py:212-397; RBFs :71-86) for a whole batch of
// proteins -- the median 3.2 s / pair of Python that feeds the hot path, and the producer of its largest input (he: 110 of
// the 123 MB a 256-pair batch ships over PCIe otherwise).
//   graph_align_frames_kernel : per protein, Kabsch-align the unbound C-alpha trace onto the bound one (:279-308), local
//                               frames n_i, u_i, v_i (:245-249) rotated along.
//   graph_knn_kernel          : per destination residue, mean all-atom distance to every residue of its protein (:324-329),
//                               neighbours = those closer than `cutoff`, or the `max_neighbor` closest (:339-343), and the
//                               5 surface-aware features mu_r_norm (:349-356).
//   graph_edges_kernel        : CSR edge list + the 27 edge features [15 RBF of the distance | p_ij, q_ij, k_ij, t_ij in the
//                               destination's local frame] (:373-389).
// Distances and features are evaluated in fp64 and rounded to fp32 once, like the reference's float64 numpy code.
// Restated in oracle/graph_oracle.py (pinned against the reference on all 125 shipped test pairs).
#include "common.cuh"
#include "svd3.cuh"

namespace eqd {

#define GB_THREADS 128
#define GB_MAXK 16
#define GB_MAXA 64     // atoms of the destination residue cached in shared memory

__global__ void __launch_bounds__(GB_THREADS)
graph_align_frames_kernel(int n_prot, const int* __restrict__ seg_ptr, const float* __restrict__ nca_c /*[n][9]*/,
                          const float* __restrict__ bound_ca /*[n][3]*/, double* __restrict__ xa /*[n][3]*/,
                          double* __restrict__ frames /*[n][9] = n, u, v*/, float* __restrict__ x32 /*[n][3]*/) {
  __shared__ double red[GB_THREADS][9];
  __shared__ double Rt[12];
  const int s = blockIdx.x, tid = threadIdx.x;
  const int i0 = seg_ptr[s], i1 = seg_ptr[s + 1], n = i1 - i0;
  auto bsum = [&](double* v, int m) {
    __syncthreads();
    for (int q = 0; q < m; ++q) red[tid][q] = v[q];
    __syncthreads();
    for (int st = GB_THREADS / 2; st > 0; st >>= 1) {
      if (tid < st)
        for (int q = 0; q < m; ++q) red[tid][q] += red[tid + st][q];
      __syncthreads();
    }
    for (int q = 0; q < m; ++q) v[q] = red[0][q];
    __syncthreads();
  };
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = i0 + tid; i < i1; i += GB_THREADS)
    for (int q = 0; q < 3; ++q) {
      c[q] += (double)nca_c[(long)i * 9 + 3 + q];
      c[3 + q] += (double)bound_ca[(long)i * 3 + q];
    }
  bsum(c, 6);
  for (int q = 0; q < 6; ++q) c[q] /= (double)(n > 0 ? n : 1);
  double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = i0 + tid; i < i1; i += GB_THREADS)
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q)
        h[r * 3 + q] += ((double)nca_c[(long)i * 9 + 3 + r] - c[r]) * ((double)bound_ca[(long)i * 3 + q] - c[3 + q]);
  bsum(h, 9);
  if (tid == 0) {
    double U[9], S[3], V[9], R[9];
    svd3(h, U, S, V);
    for (int r = 0; r < 3; ++r)
      for (int q = 0; q < 3; ++q) R[r * 3 + q] = V[r * 3] * U[q * 3] + V[r * 3 + 1] * U[q * 3 + 1] + V[r * 3 + 2] * U[q * 3 + 2];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0.0)
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) R[r * 3 + q] = V[r * 3] * U[q * 3] + V[r * 3 + 1] * U[q * 3 + 1] - V[r * 3 + 2] * U[q * 3 + 2];
    for (int q = 0; q < 9; ++q) Rt[q] = R[q];
    for (int r = 0; r < 3; ++r) Rt[9 + r] = c[3 + r] - (R[r * 3] * c[0] + R[r * 3 + 1] * c[1] + R[r * 3 + 2] * c[2]);
  }
  __syncthreads();
  for (int i = i0 + tid; i < i1; i += GB_THREADS) {
    const float* p = nca_c + (long)i * 9;
    // local frame in fp32 like the reference (:245-249), then rotated in fp64
    float ux = p[0] - p[3], uy = p[1] - p[4], uz = p[2] - p[5];
    float tx = p[6] - p[3], ty = p[7] - p[4], tz = p[8] - p[5];
    float nu = sqrtf(ux * ux + uy * uy + uz * uz), nt = sqrtf(tx * tx + ty * ty + tz * tz);
    ux /= nu; uy /= nu; uz /= nu;
    tx /= nt; ty /= nt; tz /= nt;
    float nx = uy * tz - uz * ty, ny = uz * tx - ux * tz, nz = ux * ty - uy * tx;
    const float nn = sqrtf(nx * nx + ny * ny + nz * nz);
    nx /= nn; ny /= nn; nz /= nn;
    const float vx = ny * uz - nz * uy, vy = nz * ux - nx * uz, vz = nx * uy - ny * ux;
    const double f[9] = {nx, ny, nz, ux, uy, uz, vx, vy, vz};
    for (int a = 0; a < 3; ++a)
      for (int r = 0; r < 3; ++r)
        frames[(long)i * 9 + a * 3 + r] = Rt[r * 3] * f[a * 3] + Rt[r * 3 + 1] * f[a * 3 + 1] + Rt[r * 3 + 2] * f[a * 3 + 2];
    for (int r = 0; r < 3; ++r) {
      const double v = Rt[r * 3] * (double)p[3] + Rt[r * 3 + 1] * (double)p[4] + Rt[r * 3 + 2] * (double)p[5] + Rt[9 + r];
      xa[(long)i * 3 + r] = v;
      x32[(long)i * 3 + r] = (float)v;
    }
  }
}

// Centroid and mean atom-to-centroid distance of every residue: the bounds of the neighbour search below.
__global__ void graph_centroid_kernel(int n_nodes, const int* __restrict__ atom_ptr, const float* __restrict__ atoms,
                                      double* __restrict__ cen /*[n][4] = cx, cy, cz, rho*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int a0 = atom_ptr[i], na = atom_ptr[i + 1] - a0;
  double cx = 0.0, cy = 0.0, cz = 0.0;
  for (int a = 0; a < na; ++a) {
    cx += (double)atoms[(long)(a0 + a) * 3];
    cy += (double)atoms[(long)(a0 + a) * 3 + 1];
    cz += (double)atoms[(long)(a0 + a) * 3 + 2];
  }
  const double inv = na > 0 ? 1.0 / (double)na : 0.0;
  cx *= inv; cy *= inv; cz *= inv;
  double rho = 0.0;
  for (int a = 0; a < na; ++a) {
    const double dx = (double)atoms[(long)(a0 + a) * 3] - cx, dy = (double)atoms[(long)(a0 + a) * 3 + 1] - cy,
                 dz = (double)atoms[(long)(a0 + a) * 3 + 2] - cz;
    rho += sqrt(dx * dx + dy * dy + dz * dz);
  }
  cen[(long)i * 4] = cx; cen[(long)i * 4 + 1] = cy; cen[(long)i * 4 + 2] = cz; cen[(long)i * 4 + 3] = rho * inv;
}

// Quarter `part` of the sum over all atom pairs of |a - b|: atoms part, part + 4, ... of residue A (the LOWER-index residue of
// the pair, so that d(i,j) == d(j,i) bitwise) against every atom of B.  The four quarters are added as (s0 + s1) + (s2 + s3).
__device__ __forceinline__ double atom_distance_sum_part(const float* __restrict__ A, int na, const float* __restrict__ Bp, int nb, int part) {
  double s = 0.0;
  for (int a = part; a < na; a += 4) {
    const double ax = A[a * 3], ay = A[a * 3 + 1], az = A[a * 3 + 2];
    for (int b = 0; b < nb; ++b) {
      const double dx = ax - (double)Bp[b * 3], dy = ay - (double)Bp[b * 3 + 1], dz = az - (double)Bp[b * 3 + 2];
      s += sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  return s;
}

// One CTA per destination residue i.  Dynamic smem: drow[n] (lower bound, then the exact distance of the candidates, +inf
// otherwise) and ub[n] (upper bound; later reused as the candidate list).
//   The mean all-atom distance D_ij (:324-329, 64 fp64 square roots per residue pair) is only evaluated for residues that can
//   be among the neighbours.  With c = residue centroid and rho = mean atom-to-centroid distance,
//     |c_i - c_j|  <=  D_ij  <=  |c_i - c_j| + rho_i + rho_j        (convexity of the norm; triangle inequality)
//   so when more than K residues are CERTAINLY inside the cutoff (upper bound < cutoff) the K nearest all lie below any
//   tau that at least K upper bounds stay under (taken from a 64-bin histogram of the upper bounds: within cutoff / 64 of
//   the K-th smallest), and a residue whose lower bound exceeds tau cannot be selected; otherwise every residue whose
//   lower bound is inside the cutoff is evaluated.  Selection and output order are those of the reference:
//   the K closest in ascending distance (:342-343), or all residues inside the cutoff in ascending index (:340).
__global__ void __launch_bounds__(GB_THREADS)
graph_knn_kernel(int n_nodes, const int* __restrict__ node_seg, const int* __restrict__ seg_ptr, const int* __restrict__ atom_ptr,
                 const float* __restrict__ atoms, const double* __restrict__ xa, const double* __restrict__ cen, double cutoff,
                 int max_neighbor, int* __restrict__ deg, int* __restrict__ nbr /*[n][GB_MAXK]*/,
                 double* __restrict__ nbr_d /*[n][GB_MAXK]*/, float* __restrict__ mu_r_norm /*[n][5]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (i >= n_nodes) return;
  const int s = node_seg[i], j0 = seg_ptr[s], j1 = seg_ptr[s + 1], n = j1 - j0, ii = i - j0;
  double* drow = reinterpret_cast<double*>(smem_raw);
  double* ub = drow + n;
  int* candl = reinterpret_cast<int*>(ub);       // the candidate list reuses the upper-bound row once tau is known
  __shared__ float ai[GB_MAXA * 3];
  __shared__ int sel[GB_MAXK];
  __shared__ double seld[GB_MAXK];
  __shared__ int n_certain, n_cand, n_valid;
  __shared__ int hist[64];
  __shared__ double tau_s;
  const int a0 = atom_ptr[i], na = atom_ptr[i + 1] - a0;
  const bool cached = na <= GB_MAXA;
  if (cached)
    for (int o = tid; o < na * 3; o += GB_THREADS) ai[o] = atoms[(long)a0 * 3 + o];
  if (tid == 0) { n_certain = 0; n_cand = 0; n_valid = 0; tau_s = INFINITY; }
  if (tid < 64) hist[tid] = 0;
  const double bin_w = cutoff * (1.0 / 64.0);
  const double cix = cen[(long)i * 4], ciy = cen[(long)i * 4 + 1], ciz = cen[(long)i * 4 + 2], rhoi = cen[(long)i * 4 + 3];
  __syncthreads();
  const float* Ai = cached ? ai : atoms + (long)a0 * 3;
  // ---- bounds --------------------------------------------------------------------------------------------------------
  {
    int c = 0;
    for (int jj = tid; jj < n; jj += GB_THREADS) {
      double lo = INFINITY, hi = INFINITY;
      if (jj != ii) {
        const double* cj = cen + (long)(j0 + jj) * 4;
        const double dx = cix - cj[0], dy = ciy - cj[1], dz = ciz - cj[2];
        lo = sqrt(dx * dx + dy * dy + dz * dz);
        hi = lo + rhoi + cj[3];
        if (hi < cutoff) {
          ++c;
          atomicAdd(&hist[min(63, (int)(hi / bin_w))], 1);
        }
      }
      drow[jj] = lo;
      ub[jj] = hi;
    }
    if (c) atomicAdd(&n_certain, c);
  }
  __syncthreads();
  const bool certain = n_certain > max_neighbor;     // more than K residues are inside the cutoff whatever their exact distance
  if (certain && tid < 32) {   // tau = upper edge of the first histogram bin at which the running count reaches K
    const int h0 = hist[2 * lane], h1 = hist[2 * lane + 1];
    int run = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, run, o);
      if (lane >= o) run += v;
    }
    const int before = run - h0 - h1;    // elements in the bins below 2 * lane
    int b = 64;
    if (before < max_neighbor && before + h0 >= max_neighbor) b = 2 * lane;
    else if (before + h0 < max_neighbor && run >= max_neighbor) b = 2 * lane + 1;
    if (b < 64) tau_s = fmin(cutoff, (double)(b + 1) * bin_w * (1.0 + 1e-12));
  }
  __syncthreads();
  const double thr = certain ? tau_s : cutoff;
  const double thr_safe = thr + 1e-9 * (1.0 + thr);   // the bounds hold exactly in real arithmetic; fp64 rounding is ~1e-15
  // ---- candidates -----------------------------------------------------------------------------------------------------
  for (int jj = tid; jj < n; jj += GB_THREADS) {   // (the upper bounds are dead since the barrier above: the list may overwrite them)
    if (jj != ii && drow[jj] <= thr_safe) candl[atomicAdd(&n_cand, 1)] = jj;
    else drow[jj] = INFINITY;
  }
  __syncthreads();
  const int ncand = n_cand;
  // ---- exact distances: 4 lanes per candidate ------------------------------------------------------------------------------
  for (int t0 = tid - lane; t0 < ncand * 4; t0 += GB_THREADS) {
    const int t = t0 + lane;
    const bool act = t < ncand * 4;
    double part_sum = 0.0;
    int jj = 0, nb = 1;
    if (act) {
      jj = candl[t >> 2];
      const int j = j0 + jj, b0 = atom_ptr[j];
      nb = atom_ptr[j + 1] - b0;
      const float* Bj = atoms + (long)b0 * 3;
      part_sum = j > i ? atom_distance_sum_part(Ai, na, Bj, nb, t & 3) : atom_distance_sum_part(Bj, nb, Ai, na, t & 3);
    }
    const double o1 = __shfl_xor_sync(0xffffffffu, part_sum, 1);
    const double s01 = (lane & 1) ? o1 + part_sum : part_sum + o1;        // (s0 + s1) resp. (s2 + s3), same operand order in both lanes
    const double o2 = __shfl_xor_sync(0xffffffffu, s01, 2);
    const double tot = (lane & 2) ? o2 + s01 : s01 + o2;                  // (s0 + s1) + (s2 + s3)
    if (act && (t & 3) == 0) {
      const double d = tot / ((double)na * (double)nb);
      drow[jj] = d;
      if (d < cutoff) atomicAdd(&n_valid, 1);
    }
  }
  __syncthreads();
  const int nvalid = certain ? max_neighbor + 1 : n_valid;
  const bool by_distance = nvalid > max_neighbor;       // (:342-343) the max_neighbor closest, ascending distance;
  const int k_out = by_distance ? max_neighbor : nvalid;  // otherwise every residue within the cutoff, ascending index (:340)
  // ---- selection: the output slot of a candidate is its rank ---------------------------------------------------------------
  for (int c = tid; c < ncand; c += GB_THREADS) {
    const int jj = candl[c];
    const double d = drow[jj];
    if (!(d < cutoff)) continue;
    int rank = 0;
    for (int q = 0; q < ncand; ++q) {
      const int jq = candl[q];
      const double dq = drow[jq];
      if (!(dq < cutoff)) continue;
      rank += by_distance ? ((dq < d || (dq == d && jq < jj)) ? 1 : 0) : (jq < jj ? 1 : 0);
    }
    if (rank < k_out) {
      sel[rank] = j0 + jj;
      seld[rank] = d;
    }
  }
  __syncthreads();
  if (tid < k_out) {
    nbr[(long)i * GB_MAXK + tid] = sel[tid];
    nbr_d[(long)i * GB_MAXK + tid] = seld[tid];
  }
  if (tid == 0) deg[i] = k_out;
  if (tid < 5) {                     // mu_r_norm (:349-356): softmax over the neighbours of -d^2 / sigma
    const double sig = tid == 0 ? 1.0 : (tid == 1 ? 2.0 : (tid == 2 ? 5.0 : (tid == 3 ? 10.0 : 30.0)));
    double mx = -INFINITY;
    for (int k = 0; k < k_out; ++k) mx = fmax(mx, -seld[k] * seld[k] / sig);
    double wsum = 0.0, mvx = 0.0, mvy = 0.0, mvz = 0.0, den = 0.0;
    for (int k = 0; k < k_out; ++k) {
      const double w = exp(-seld[k] * seld[k] / sig - mx);
      const int j = sel[k];
      const double dx = xa[(long)i * 3] - xa[(long)j * 3], dy = xa[(long)i * 3 + 1] - xa[(long)j * 3 + 1],
                   dz = xa[(long)i * 3 + 2] - xa[(long)j * 3 + 2];
      wsum += w;
      mvx += w * dx; mvy += w * dy; mvz += w * dz;
      den += w * sqrt(dx * dx + dy * dy + dz * dz);
    }
    mu_r_norm[(long)i * 5 + tid] = k_out > 0 ? (float)(sqrt(mvx * mvx + mvy * mvy + mvz * mvz) / den) : 0.f;
  }
}

// thread per (destination node, neighbour slot)
__global__ void graph_edges_kernel(int n_nodes, const int* __restrict__ row_ptr, const int* __restrict__ deg,
                                   const int* __restrict__ nbr, const double* __restrict__ nbr_d,
                                   const double* __restrict__ xa, const double* __restrict__ frames,
                                   int* __restrict__ col_src, int* __restrict__ edge_dst, float* __restrict__ he) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(t / GB_MAXK), k = (int)(t - (long)i * GB_MAXK);
  if (i >= n_nodes || k >= deg[i]) return;
  const int e = row_ptr[i] + k, j = nbr[(long)i * GB_MAXK + k];
  col_src[e] = j;
  edge_dst[e] = i;
  float* o = he + (long)e * EQD_EDGE_FEATS;
  const double d = nbr_d[(long)i * GB_MAXK + k], d2 = d * d;
  double sigma = 1.0;
  for (int q = 0; q < 15; ++q) {                 // distance_list_featurizer (:71-86)
    o[q] = (float)exp(-d2 / sigma);
    sigma *= 1.5;
  }
  const double* B = frames + (long)i * 9;         // rows n_i, u_i, v_i of the DESTINATION (:378)
  const double* Fj = frames + (long)j * 9;
  const double rel[3] = {xa[(long)j * 3] - xa[(long)i * 3], xa[(long)j * 3 + 1] - xa[(long)i * 3 + 1], xa[(long)j * 3 + 2] - xa[(long)i * 3 + 2]};
  for (int r = 0; r < 3; ++r) {
    o[15 + r] = (float)(B[r * 3] * rel[0] + B[r * 3 + 1] * rel[1] + B[r * 3 + 2] * rel[2]);          // p_ij
    for (int a = 0; a < 3; ++a)                                                                       // q_ij, k_ij, t_ij
      o[18 + a * 3 + r] = (float)(B[r * 3] * Fj[a * 3] + B[r * 3 + 1] * Fj[a * 3 + 1] + B[r * 3 + 2] * Fj[a * 3 + 2]);
  }
}

__global__ void graph_node_seg_kernel(int n_prot, const int* __restrict__ seg_ptr, int* __restrict__ node_seg) {
  const int s = blockIdx.x;
  for (int i = seg_ptr[s] + threadIdx.x; i < seg_ptr[s + 1]; i += blockDim.x) node_seg[i] = s;
}

}  // namespace eqd

extern "C" size_t eqd_graph_build_workspace_bytes(int32_t n_nodes) {
  const size_t N = n_nodes > 0 ? n_nodes : 1;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  return al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4) + al(N * GB_MAXK * 4) + al(N * GB_MAXK * 8) + al(N * 4 * 8);
}

// Stage 1: alignment, frames, neighbour search.  Outputs deg [n] (the caller turns it into row_ptr with an exclusive
// prefix sum), x [n][3] fp32 (ndata['x']), mu_r_norm [n][5].  max_protein_nodes sizes the per-CTA distance row.
extern "C" int eqd_graph_build_knn(int32_t n_prot, int32_t n_nodes, int32_t max_protein_nodes, const int32_t* seg_ptr,
                                   const int32_t* atom_ptr, const float* atoms, const float* nca_c, const float* bound_ca,
                                   float cutoff, int32_t max_neighbor, void* workspace, size_t workspace_bytes, int32_t* deg,
                                   float* x, float* mu_r_norm, void* stream) {
  if (!seg_ptr || !atom_ptr || !atoms || !nca_c || !bound_ca || !workspace || !deg || !x || !mu_r_norm) return EQD_ERR_BAD_ARG;
  if (max_neighbor < 1 || max_neighbor > GB_MAXK || max_protein_nodes < 1) return EQD_ERR_UNSUPPORTED;
  if (workspace_bytes < eqd_graph_build_workspace_bytes(n_nodes)) return EQD_ERR_WORKSPACE;
  if (n_nodes <= 0 || n_prot <= 0) return EQD_OK;
  const size_t row_bytes = (size_t)max_protein_nodes * 2 * sizeof(double);   // lower-bound / distance row + upper-bound row
  if (row_bytes > 200 * 1024) return EQD_ERR_UNSUPPORTED;      // > 12800 residues in one protein
  cudaStream_t st = (cudaStream_t)stream;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t N = n_nodes;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  double* xa = reinterpret_cast<double*>(w);
  double* frames = reinterpret_cast<double*>(w + al(N * 3 * 8));
  int* node_seg = reinterpret_cast<int*>(w + al(N * 3 * 8) + al(N * 9 * 8));
  int* nbr = reinterpret_cast<int*>(w + al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4));
  double* nbr_d = reinterpret_cast<double*>(w + al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4) + al(N * GB_MAXK * 4));
  double* cen = reinterpret_cast<double*>(w + al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4) + al(N * GB_MAXK * 4) + al(N * GB_MAXK * 8));
  eqd::graph_centroid_kernel<<<(n_nodes + 127) / 128, 128, 0, st>>>(n_nodes, atom_ptr, atoms, cen);
  EQD_CUDA_LAUNCH_CHECK();
  eqd::graph_align_frames_kernel<<<n_prot, GB_THREADS, 0, st>>>(n_prot, seg_ptr, nca_c, bound_ca, xa, frames, x);
  EQD_CUDA_LAUNCH_CHECK();
  eqd::graph_node_seg_kernel<<<n_prot, 128, 0, st>>>(n_prot, seg_ptr, node_seg);
  EQD_CUDA_LAUNCH_CHECK();
  EQD_SET_SMEM((eqd::graph_knn_kernel), row_bytes);
  eqd::graph_knn_kernel<<<n_nodes, GB_THREADS, row_bytes, st>>>(n_nodes, node_seg, seg_ptr, atom_ptr, atoms, xa, cen, (double)cutoff,
                                                              max_neighbor, deg, nbr, nbr_d, mu_r_norm);
  EQD_CUDA_LAUNCH_CHECK();
  return EQD_OK;
}

// Stage 2 (after row_ptr = exclusive prefix sum of deg): CSR edge list and the (E, 27) edge features.
extern "C" int eqd_graph_build_edges(int32_t n_nodes, const int32_t* row_ptr, const int32_t* deg, const void* workspace,
                                     int32_t* col_src, int32_t* edge_dst, float* he, void* stream) {
  if (!row_ptr || !deg || !workspace || !col_src || !edge_dst || !he) return EQD_ERR_BAD_ARG;
  if (n_nodes <= 0) return EQD_OK;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t N = n_nodes;
  const unsigned char* w = reinterpret_cast<const unsigned char*>(workspace);
  const double* xa = reinterpret_cast<const double*>(w);
  const double* frames = reinterpret_cast<const double*>(w + al(N * 3 * 8));
  const int* nbr = reinterpret_cast<const int*>(w + al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4));
  const double* nbr_d = reinterpret_cast<const double*>(w + al(N * 3 * 8) + al(N * 9 * 8) + al(N * 4) + al(N * GB_MAXK * 4));
  const long threads = (long)n_nodes * GB_MAXK;
  eqd::graph_edges_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_nodes, row_ptr, deg, nbr, nbr_d, xa,
                                                                                               frames, col_src, edge_dst, he);
  EQD_CUDA_LAUNCH_CHECK();
  return EQD_OK;
}
