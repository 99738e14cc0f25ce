// Dense EAM Hessian assembly for one replica -- the B200 counterpart of potential::EAM::hessian
// (src/potential/EAM/eam.cpp:122-233) writing a system::Hessian (include/libfly/system/hessian.hpp:42-148):
// column-major (3N)x(3N) fp64, 3x3 block (i,j) at rows 3i.., columns 3j.., lower triangle only,
// frozen rows/columns left zero.
//
//   pass 1  rho_b and mu_b = -sum_a phi'_{a->b}(r) dr / r                       (eam.cpp:129-143)
//   pass 2  one warp per column block z (the reference's ownership, eam.cpp:165-232: everything pass 2 writes for z
//           lies in columns 3z..3z+2): the (z,a) pair terms are computed redundantly on all lanes (uniform loop over
//           a), the 9 elements of the diagonal block are kept in registers, the off-diagonal block (a,z) is added
//           by 9 lanes, and the neighbour-of-neighbour overlap terms (z - a - g) are spread over the 32 lanes.
//           Different paths reach the same block (g,z), but always from the SAME warp, so there are no atomics:
//           plain read-modify-writes, ordered by __syncwarp() between the steps of the loop over a; lanes that hit
//           the same block inside one step (the same real atom g through two periodic images) add one after the
//           other in lane order. The sum order is fixed, so the matrix is bit-reproducible.
// Batched: one launch assembles the Hessians of `count` replicas (selection list), matrix s at H + s * stride.
// This is assembly into a dense matrix (a scatter by nature); it runs once per successful search,
// not inside the search loop.
#pragma once

#include "eam.cuh"

namespace ofly {

  // aux layout: [N][4] = {rho, mu_x, mu_y, mu_z}
  // aux of selected replica s starts at aux + s * 4 N
  template <int G>
  static __global__ void __launch_bounds__(256) k_hess_rho_mu(BatchDev B, TablesDev T, const int *__restrict__ sel, int count,
                                                              double *__restrict__ aux_all) {
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G;
    int lane = threadIdx.x % G;
    if (g >= (size_t)B.N * count) return;
    const int s = (int)(g / B.N), r = sel[s];
    int i = (int)(g % B.N);
    double *aux = aux_all + (size_t)s * 4 * B.N;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    const uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
    const int n = B.cnt[(size_t)r * B.N + i];
    const int ti = B.type[i];
    const double4 pi = ld256(pos + i);
    bool oob = false;
    double rho = 0, mx = 0, my = 0, mz = 0;
    for (int k = lane; k < n; k += G) {
      PairGeom p = load_pair(B, pos, pi, row[row_slot(k)]);
      if (p.r2 < T.r_cut_sq) {
        double rinv = fast_rsqrt(p.r2);
        double rr = p.r2 * rinv;
        double x0;
        int idx = knot_index(rr, T.inv_dR, T.dR, T.kR, x0, oob);
        double4 kn = ld256_nc(T.phi + (size_t)(p.tj + T.n_types * ti) * T.kR + idx);
        rho += spline_f(kn, x0);
        double s = spline_fp(kn, x0) * rinv;
        mx = fma(-s, p.dx, mx), my = fma(-s, p.dy, my), mz = fma(-s, p.dz, mz);
      }
    }
    rho = group_sum<G>(rho), mx = group_sum<G>(mx), my = group_sum<G>(my), mz = group_sum<G>(mz);
    if (lane == 0) {
      aux[4 * i] = rho, aux[4 * i + 1] = mx, aux[4 * i + 2] = my, aux[4 * i + 3] = mz;
    }
    if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
  }

  __device__ __forceinline__ void embed_derivs(const TablesDev &T, int t, double rho, double &fp, double &fpp, bool &oob) {
    double x0;
    int idx = knot_index(rho, T.inv_dF, T.dF, T.kF, x0, oob);
    double4 k = ld256_nc(T.F + (size_t)t * T.kF + idx);
    fp = spline_fp(k, x0), fpp = spline_fpp(k, x0);
  }

  // H(3*bi + row, 3*bj + col) += s * u[row] * v[col] for the 9 (row, col) of a block (plain read-modify-write: the
  // calling warp owns the column block).
  __device__ __forceinline__ void add_outer(double *H, size_t n3, int bi, int bj, double s, const double u[3], const double v[3]) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int rw = 0; rw < 3; rw++) {
        double *p = H + (size_t)(3 * bj + c) * n3 + (size_t)(3 * bi + rw);
        *p = *p + s * u[rw] * v[c];
      }
    }
  }

  template <int G>
  static __global__ void __launch_bounds__(256) k_hess_columns(BatchDev B, TablesDev T, const int *__restrict__ sel, int count,
                                                               const double *__restrict__ aux_all, double *__restrict__ H_all, size_t h_stride) {
    static_assert(G == 32, "one full warp per column block");
    size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G;
    int lane = threadIdx.x % G;
    if (w >= (size_t)B.N * count) return;
    const int s = (int)(w / B.N), r = sel[s];
    int z = (int)(w % B.N);
    if (B.frozen[z]) return;
    const double *aux = aux_all + (size_t)s * 4 * B.N;
    double *H = H_all + (size_t)s * h_stride;

    const size_t n3 = 3 * (size_t)B.N;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    const uint32_t *row_z = B.nbr + ((size_t)r * B.N + z) * B.K;
    const int n_z = B.cnt[(size_t)r * B.N + z];
    const int tz = B.type[z];
    const double4 pz = ld256(pos + z);
    bool oob = false;

    double Fp_z, Fpp_z;
    embed_derivs(T, tz, aux[4 * z], Fp_z, Fpp_z, oob);
    const double mu_z[3] = {aux[4 * z + 1], aux[4 * z + 2], aux[4 * z + 3]};

    // diagonal block, kept on every lane (identical values), written by lanes 0..8 at the end
    double D[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) D[a][b] = Fpp_z * mu_z[a] * mu_z[b];

    for (int k = 0; k < n_z; k++) {
      PairGeom p = load_pair(B, pos, pz, row_z[row_slot(k)]);
      if (!(p.r2 < T.r_cut_sq)) continue;  // warp-uniform
      const int a = p.j, ta = p.tj;
      double rinv = fast_rsqrt(p.r2);
      double rr = p.r2 * rinv;
      double x0;
      int idx = knot_index(rr, T.inv_dR, T.dR, T.kR, x0, oob);
      int hi = max(tz, ta), lo = min(tz, ta);
      double4 kV = ld256_nc(T.V + (size_t)(hi + T.n_types * lo) * T.kR + idx);
      double4 kaz = ld256_nc(T.phi + (size_t)(ta + T.n_types * tz) * T.kR + idx);  // phi(a -> z)
      double4 kza = ld256_nc(T.phi + (size_t)(tz + T.n_types * ta) * T.kR + idx);  // phi(z -> a)
      double Fp_a, Fpp_a;
      embed_derivs(T, ta, aux[4 * a], Fp_a, Fpp_a, oob);

      double paz = spline_fp(kaz, x0), pza = spline_fp(kza, x0);
      double A = (spline_fp(kV, x0) + Fp_z * paz + Fp_a * pza) * rinv;                           // A.14
      double Bq = spline_fpp(kV, x0) + Fp_z * spline_fpp(kaz, x0) + Fp_a * spline_fpp(kza, x0);  // A.14
      double rinv2 = rinv * rinv;
      double ABFpp = (A - Bq - Fpp_a * (pza * pza)) * rinv2;
      const double dr[3] = {p.dx, p.dy, p.dz};
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) D[i][j] -= ABFpp * dr[i] * dr[j];
        D[i][i] += A;
      }

      if (a > z && !B.frozen[a] && lane < 9) {  // block (a, z), one element per lane
        const double mu_a[3] = {aux[4 * a + 1], aux[4 * a + 2], aux[4 * a + 3]};
        double BArr = (Bq - A) * rinv2;
        double ur = Fpp_z * paz * rinv;
        double ru = Fpp_a * pza * rinv;
        int rw = lane % 3, c = lane / 3;
        double v = -BArr * dr[rw] * dr[c] + ur * dr[rw] * mu_z[c] - ru * mu_a[rw] * dr[c];
        if (rw == c) v -= A;
        double *q = H + (size_t)(3 * z + c) * n3 + (size_t)(3 * a + rw);
        *q = *q + v;
      }
      __syncwarp();  // the direct block is in memory before an overlap term may add to it (a's row can hold a's own image)

      // overlap terms z - a - g
      double ddFg = Fpp_a * pza;
      const double4 pa_real = ld256(pos + a);
      // neighbours of a are relative to the *real* atom a: dr_ag is image-independent
      const uint32_t *row_a = B.nbr + ((size_t)r * B.N + a) * B.K;
      const int n_a = B.cnt[(size_t)r * B.N + a];
      for (int q0 = 0; q0 < n_a; q0 += G) {  // warp-uniform trip count
        const int q = q0 + lane;
        bool take = false;
        int g = -1;
        double mag = 0;
        double dag[3] = {0, 0, 0};
        if (q < n_a) {
          PairGeom pg = load_pair(B, pos, pa_real, row_a[row_slot(q)]);
          if (pg.r2 < T.r_cut_sq) {
            g = pg.j;
            take = (g > z && !B.frozen[g]);
            if (g == z) {
              double sx = p.dx + pg.dx, sy = p.dy + pg.dy, sz = p.dz + pg.dz;
              take = sqrt(fma(sz, sz, fma(sy, sy, sx * sx))) > 1e-6;  // a periodic image of z itself (eam.cpp:223)
            }
            if (take) {
              double rinv_ag = fast_rsqrt(pg.r2);
              double r_ag = pg.r2 * rinv_ag;
              double y0;
              int idg = knot_index(r_ag, T.inv_dR, T.dR, T.kR, y0, oob);
              double pga = spline_fp(ld256_nc(T.phi + (size_t)(pg.tj + T.n_types * ta) * T.kR + idg), y0);  // phi(g -> a)
              mag = ddFg * rinv * rinv_ag * pga;
              dag[0] = pg.dx, dag[1] = pg.dy, dag[2] = pg.dz;
            }
          }
        }
        // lanes that target the same block (g reached through two images) take turns in lane order
        const unsigned peers = __match_any_sync(0xffffffffu, take ? g : -1 - lane);
        const int turn = __popc(peers & ((1u << lane) - 1u));
        const int turns = __reduce_max_sync(0xffffffffu, take ? __popc(peers) : 0);
        for (int t = 0; t < turns; t++) {
          if (take && turn == t) add_outer(H, n3, g, z, -mag, dag, dr);
          __syncwarp();
        }
      }
    }
    if (lane < 9) {
      int rw = lane % 3, c = lane / 3;
      double v = 0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
          if (i == rw && j == c) v = D[i][j];
      double *q = H + (size_t)(3 * z + c) * n3 + (size_t)(3 * z + rw);
      *q = *q + v;
    }
    if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
  }

  // Hessian::mass_weight (hessian.hpp:136-143): every element scaled by 1/sqrt(m_row m_col); `count` matrices, h_stride apart.
  static __global__ void k_hess_mass_weight(double *__restrict__ H_all, size_t h_stride, int count, int N, const uint8_t *__restrict__ type,
                                            const double *__restrict__ mass) {
    size_t n3 = 3 * (size_t)N;
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= n3 * n3) return;
    size_t col = t / n3, rw = t % n3;
    double mi = mass[type[col / 3]];
    const double f = 1 / sqrt(mi * mass[type[rw / 3]]);
    for (int s = 0; s < count; s++) H_all[(size_t)s * h_stride + t] *= f;
  }

}  // namespace ofly
