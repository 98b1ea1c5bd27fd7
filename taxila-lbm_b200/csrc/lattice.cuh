// lattice.cuh -- compile-time lattice descriptions for the device code.
//
// The numbers mirror the reference's DiscretizationSetup_D3Q19
// (src/lbm/lbm_discretization_d3q19.F90:64-232) and DiscretizationSetUp_D2Q9
// (src/lbm/lbm_discretization_d2q9.F90:53-144): direction order, lattice vectors,
// weights, MRT moment rows and their squared norms, c_0, and the isotropy weights
// ffw(L).  Everything is constexpr so that the unrolled kernels fold the tables into
// immediates; opposite directions are derived from the vectors (the reference lists
// them, :83-161, and they are exactly the negated vectors).
#pragma once
#include <type_traits>

#include "ff_stencil.cuh"

#define TXG_HD __host__ __device__ __forceinline__

namespace txg {

// compile-time loop: f(std::integral_constant<int, I>) for I in [B, E)
template <int B, int E, class F>
TXG_HD constexpr void static_for(F &&f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

struct D3Q19 {
  static constexpr int D = 3, Q = 19;
  static constexpr int NAXIS = 6;   // directions 1..6 are the axis neighbours
  static constexpr int NCROSS = 5;  // directions with c_z = +1 (and as many with c_z = -1)
  using FF = FFStencilD3;
  TXG_HD static constexpr int c(int n, int d) {
    // lbm_discretization_d3q19.F90:163-168
    constexpr int t[19][3] = {{0, 0, 0},  {1, 0, 0},  {0, 1, 0},   {-1, 0, 0}, {0, -1, 0}, {0, 0, 1},  {0, 0, -1},
                              {1, 1, 0},  {-1, 1, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 0, 1},  {-1, 0, 1}, {-1, 0, -1},
                              {1, 0, -1}, {0, 1, 1},  {0, -1, 1},  {0, -1, -1}, {0, 1, -1}};
    return t[n][d];
  }
  TXG_HD static constexpr double w(int n) {  // :172-175
    return n == 0 ? 1.0 / 3.0 : (n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0);
  }
  TXG_HD static constexpr int M(int r, int i) {  // rows of the moment matrix, :177-195
    constexpr int t[19][19] = {
        {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
        {-30, -11, -11, -11, -11, -11, -11, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8},
        {12, -4, -4, -4, -4, -4, -4, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1},
        {0, 1, 0, -1, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 0, 0, 0, 0},
        {0, -4, 0, 4, 0, 0, 0, 1, -1, -1, 1, 1, -1, -1, 1, 0, 0, 0, 0},
        {0, 0, 1, 0, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, -1, 1},
        {0, 0, -4, 0, 4, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, -1, 1},
        {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
        {0, 0, 0, 0, 0, -4, 4, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1},
        {0, 2, -1, 2, -1, -1, -1, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
        {0, -4, 2, -4, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, -2, -2, -2, -2},
        {0, 0, 1, 0, 1, -1, -1, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
        {0, 0, -2, 0, -2, 2, 2, 1, 1, 1, 1, -1, -1, -1, -1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 1, -1, -1, 1, -1, 1, 1, -1, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, -1, -1, 1, 1, 0, 0, 0, 0, 1, -1, -1, 1},
        {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, -1, -1, -1, -1, 1, 1}};
    return t[r][i];
  }
  TXG_HD static constexpr int Mnorm(int r) {  // :197
    constexpr int t[19] = {19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8};
    return t[r];
  }
  // which of the seven MRT rates relaxes moment r (:243-261): 0 s_c 1 s_e 2 s_e2 3 s_q 4 s_nu 5 s_pi 6 s_m
  TXG_HD static constexpr int rate_of(int r) {
    constexpr int t[19] = {0, 1, 2, 0, 3, 0, 3, 0, 3, 4, 5, 4, 5, 4, 4, 4, 6, 6, 6};
    return t[r];
  }
  // isotropy weights ffw(L) (:199-213); D3 order 10 is an LBMError in the reference (:216-229)
  TXG_HD static constexpr double ffw(int order, int L) {
    if (order == 4) return L == 1 ? 1.0 / 6.0 : (L == 2 ? 1.0 / 12.0 : 0.0);
    if (order == 8) {
      constexpr double t[9] = {0.0,         4.0 / 45.0,  1.0 / 21.0, 2.0 / 105.0, 5.0 / 504.0,
                               1.0 / 315.0, 1.0 / 630.0, 0.0,        1.0 / 5040.0};
      return L <= 8 ? t[L] : 0.0;
    }
    return 0.0;
  }
  TXG_HD static constexpr bool order_ok(int order) { return order == 4 || order == 8; }
  // feq_0 / rho (lbm_discretization_d3q19.F90:290)
  TXG_HD static double feq0(double d_k, double usqr) { return d_k - usqr / 2.; }
  // fluid-solid weights are default-real literals 1./6. and 1./12. (lbm_forcing.F90:1355,1364)
  TXG_HD static constexpr double fs_weight(int n) { return n <= 6 ? (double)(1.f / 6.f) : (double)(1.f / 12.f); }
};

struct D2Q9 {
  static constexpr int D = 2, Q = 9;
  static constexpr int NAXIS = 4;
  static constexpr int NCROSS = 0;
  using FF = FFStencilD2;
  TXG_HD static constexpr int c(int n, int d) {
    // lbm_discretization_d2q9.F90:99-100
    constexpr int t[9][3] = {{0, 0, 0},  {1, 0, 0},  {0, 1, 0},   {-1, 0, 0}, {0, -1, 0},
                             {1, 1, 0},  {-1, 1, 0}, {-1, -1, 0}, {1, -1, 0}};
    return t[n][d];
  }
  TXG_HD static constexpr double w(int n) {  // :104-106
    return n == 0 ? 4.0 / 9.0 : (n <= 4 ? 1.0 / 9.0 : 1.0 / 36.0);
  }
  TXG_HD static constexpr int M(int r, int i) {  // :108-116
    constexpr int t[9][9] = {{1, 1, 1, 1, 1, 1, 1, 1, 1},     {-4, -1, -1, -1, -1, 2, 2, 2, 2},
                             {4, -2, -2, -2, -2, 1, 1, 1, 1}, {0, 1, 0, -1, 0, 1, -1, -1, 1},
                             {0, -2, 0, 2, 0, 1, -1, -1, 1},  {0, 0, 1, 0, -1, 1, 1, -1, -1},
                             {0, 0, -2, 0, 2, 1, 1, -1, -1},  {0, 1, -1, 1, -1, 0, 0, 0, 0},
                             {0, 0, 0, 0, 0, 1, -1, 1, -1}};
    return t[r][i];
  }
  TXG_HD static constexpr int Mnorm(int r) {  // :118
    constexpr int t[9] = {9, 36, 36, 6, 12, 6, 12, 4, 4};
    return t[r];
  }
  TXG_HD static constexpr int rate_of(int r) {  // :155-163
    constexpr int t[9] = {0, 1, 2, 0, 3, 0, 3, 4, 4};
    return t[r];
  }
  TXG_HD static constexpr double ffw(int order, int L) {  // :120-142
    if (order == 4) return L == 1 ? 1.0 / 3.0 : (L == 2 ? 1.0 / 12.0 : 0.0);
    if (order == 8) {
      constexpr double t[9] = {0.0, 4.0 / 21.0, 4.0 / 45.0, 0.0, 1.0 / 60.0, 2.0 / 315.0, 0.0, 0.0, 1.0 / 5040.0};
      return L <= 8 ? t[L] : 0.0;
    }
    if (order == 10) {
      constexpr double t[11] = {0.0,         262.0 / 1785.0, 93.0 / 1190.0, 0.0,          7.0 / 340.0, 6.0 / 595.0,
                                0.0,         0.0,            9.0 / 9520.0,  2.0 / 5355.0, 1.0 / 7140.0};
      return L <= 10 ? t[L] : 0.0;
    }
    return 0.0;
  }
  TXG_HD static constexpr bool order_ok(int order) { return order == 4 || order == 8 || order == 10; }
  // lbm_discretization_d2q9.F90:193
  TXG_HD static double feq0(double d_k, double usqr) { return (1. + d_k * 5.) / 6. - 2. * usqr / 3.; }
  // 1./3. and 1./12. default-real literals (lbm_forcing.F90:1406,1414)
  TXG_HD static constexpr double fs_weight(int n) { return n <= 4 ? (double)(1.f / 3.f) : (double)(1.f / 12.f); }
};

// direction index of the vector (x,y,z), or -1
template <class L>
TXG_HD constexpr int dir_of(int x, int y, int z) {
  for (int n = 0; n < L::Q; ++n)
    if (L::c(n, 0) == x && L::c(n, 1) == y && L::c(n, 2) == z) return n;
  return -1;
}
template <class L>
TXG_HD constexpr int opp(int n) {
  return dir_of<L>(-L::c(n, 0), -L::c(n, 1), -L::c(n, 2));
}
// Centre directions: the lattice directions with c_x = 0 (n >= 1).  The adjacency table stores the
// position of X + c_n for these only; the c_x = +-1 neighbours follow from them and the mask bits,
// because positions run along x (hot_kernels.cuh, Adjacency).
template <class L>
TXG_HD constexpr int num_centres() {
  int k = 0;
  for (int n = 1; n < L::Q; ++n)
    if (L::c(n, 0) == 0) ++k;
  return k;
}
// rank of direction n among the centre directions, or -1
template <class L>
TXG_HD constexpr int centre_rank(int n) {
  if (n < 1 || L::c(n, 0) != 0) return -1;
  int k = 0;
  for (int i = 1; i < n; ++i)
    if (L::c(i, 0) == 0) ++k;
  return k;
}
// number of fluid-fluid stencil entries used at a given isotropy order
template <class L>
TXG_HD constexpr int ff_entries(int order) {
  int k = 0;
  for (int e = 0; e < L::FF::E; ++e)
    if (L::FF::gate[e] <= order) ++k;
  return k;
}
template <class L>
TXG_HD constexpr int ff_words(int order) {
  return (ff_entries<L>(order) + 31) / 32;
}
// ghost width the stencil needs (lbm_grid.F90:107-120)
TXG_HD constexpr int stencil_radius(int order) { return order == 4 ? 1 : (order == 8 ? 2 : (order == 10 ? 3 : 1)); }

}  // namespace txg
