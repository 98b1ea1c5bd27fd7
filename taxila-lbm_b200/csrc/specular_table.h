// specular_table.h -- host-only (plain C++, no CUDA): which population slots the free-slip walls
// (WALL_NORMAL_X/Y/Z = 900-902, device classes 250-252) rewrite after every push, from the geometry alone.
// Included by flow.cu (txg_set_walls) and compiled on its own by tests/test_specular_table.py, which replays
// the push + this table in numpy against the oracle's literal stream + bounce-back sweep.
//
// DistributionBouncebackD3/D2 (lbm_distribution_function.F90:669-784) treats a WALL_NORMAL_a node W as a
// mirror: for every n, fi(reflect_a(n), W - c_n,a e_a) = fi(n, W), i.e. the population that left the fluid
// node S = W - c_n along c_n arrives with its normal component reversed at T = S + (tangential part of
// c_n).  The push kernels know only plain walls (they park that population in slot (opp(n), S)).  For every
// fluid node T next to a 900-902 node and every direction nn this table says what the reference's sweep
// leaves in fi(nn, T): the candidates are the stream from A = T - c_nn, the plain bounce-back off A, and
// one mirror W_a = T - c_nn,a e_a per axis; they write in the sweep's (k, j, i) order and the last wins.
//
// `parked` counts reflections off a mirror that would land on a solid node (free-slip walls meeting in
// a corner, an obstacle touching the wall).  The reference parks such a population in the wall node's own
// storage and its fate depends on the sweep order; the fluid-only device storage has no such slot, so the
// caller refuses the geometry (PETSC_ERR_SUP) instead of computing something else.
#pragma once
#include <cstdint>
#include <vector>

namespace txg {

struct LatticeTab {
  int Q, D;
  int c[19][3];
  double w[19];
};

constexpr uint32_t SPEC_ZERO = 0xffffffffu;

struct SpecularTable {
  std::vector<uint32_t> dst, src;  // element offsets n*fs + pos inside one component's block; src may be SPEC_ZERO
  long long parked = 0;
};

// cls: [NZl+2Rz][NY+2R][NX+2R] node classes with ghosts (periodic images / 255 outside), R >= 1
// P:   [nE+1] extended node index -> position, or nullptr (identity: no solid node at all)
inline void build_specular_table(const LatticeTab &lt, int NX, int NY, int NZl, int R, int Rz, const int per[3],
                                 const uint8_t *cls, const uint32_t *P, long long fs, SpecularTable &out) {
  const int Q = lt.Q, D = lt.D;
  const int cnx = NX + 2 * R, cny = NY + 2 * R;
  const long long plane = (long long)NX * NY;
  const int N[3] = {NX, NY, NZl};
  auto cls_at = [&](const int x[3]) -> int {  // class of an owned node or of a ghost node next to one
    return cls[(size_t)(((long long)(x[2] + Rz) * cny + (x[1] + R)) * cnx + (x[0] + R))];
  };
  auto pos_at = [&](const int x[3]) -> long long {  // position of a fluid node given by unwrapped coordinates
    int w[3];
    for (int d = 0; d < 3; ++d) {
      w[d] = x[d];
      if (d < D && per[d]) w[d] = ((x[d] % N[d]) + N[d]) % N[d];
    }
    const long long oe = (long long)(w[2] + Rz) * plane + (long long)w[1] * NX + w[0];
    return P ? (long long)P[(size_t)oe] : oe;
  };
  auto is_mirror = [&](int c) { return c >= 250 && c <= 252 && c - 250 < D; };  // 902 in 2-D is a plain wall
  int opp[19], refl[3][19];
  auto find = [&](int a, int b, int c) {
    for (int n = 0; n < Q; ++n)
      if (lt.c[n][0] == a && lt.c[n][1] == b && lt.c[n][2] == c) return n;
    return -1;
  };
  for (int n = 0; n < Q; ++n) {
    const int *c = lt.c[n];
    opp[n] = find(-c[0], -c[1], -c[2]);
    refl[0][n] = find(-c[0], c[1], c[2]);
    refl[1][n] = find(c[0], -c[1], c[2]);
    refl[2][n] = find(c[0], c[1], -c[2]);
  }
  // rank of a node in the bounce-back sweep over the ghosted (width 1) box
  auto rank_of = [&](const int W[3]) { return ((long long)(W[2] + 1) * (NY + 2) + (W[1] + 1)) * (NX + 2) + (W[0] + 1); };
  out.dst.clear();
  out.src.clear();
  out.parked = 0;
  for (int z = 0; z < NZl; ++z)
    for (int y = 0; y < NY; ++y)
      for (int x = 0; x < NX; ++x) {
        const int T[3] = {x, y, z};
        if (cls_at(T) != 0) continue;
        bool near = false;
        for (int n = 1; n < Q && !near; ++n) {
          const int W[3] = {x + lt.c[n][0], y + lt.c[n][1], z + lt.c[n][2]};
          near = is_mirror(cls_at(W));
        }
        if (!near) continue;
        const long long posT = pos_at(T);
        for (int nn = 1; nn < Q; ++nn) {
          const int *c = lt.c[nn];
          {  // T pushes nn into a mirror of axis a: the reflection lands on T + tangential(c_nn)
            const int W[3] = {x + c[0], y + c[1], z + c[2]};
            const int cw = cls_at(W);
            if (is_mirror(cw) && c[cw - 250] != 0) {
              int L[3] = {W[0], W[1], W[2]};
              L[cw - 250] = T[cw - 250];
              if (cls_at(L) != 0) ++out.parked;
            }
          }
          // what the reference leaves in fi(nn, T)
          const int A[3] = {x - c[0], y - c[1], z - c[2]};
          const int ca = cls_at(A);
          enum { KEEP, ZERO, COPY } what = ca == 0 ? KEEP : ZERO;  // the stream alone (solid nodes hold 0)
          long long best = -1;                                     // sweep rank of the last writer so far
          uint32_t from = SPEC_ZERO;
          if (ca != 0 && !is_mirror(ca)) {  // plain wall at A: T's own opp(nn) comes back, where the push put it
            best = rank_of(A);
            what = KEEP;
          }
          for (int a = 0; a < D; ++a) {
            if (c[a] == 0) continue;
            int W[3] = {x, y, z};
            W[a] -= c[a];
            if (cls_at(W) != 250 + a) continue;
            const long long rk = rank_of(W);
            if (rk < best) continue;
            best = rk;
            int S[3] = {A[0], A[1], A[2]};
            S[a] = T[a];  // S = T - tangential(c_nn)
            const int nsrc = refl[a][nn];
            if (S[0] == x && S[1] == y && S[2] == z) {
              what = KEEP;  // c_nn is normal to the mirror: plain bounce-back of T's own population
            } else if (cls_at(S) == 0) {
              what = COPY;  // the push parked f*_nsrc(S) in slot (opp(nsrc), S)
              from = (uint32_t)((long long)opp[nsrc] * fs + pos_at(S));
            } else {
              what = ZERO;
            }
          }
          if (what == KEEP) continue;
          out.dst.push_back((uint32_t)((long long)nn * fs + posT));
          out.src.push_back(what == COPY ? from : SPEC_ZERO);
        }
      }
}

}  // namespace txg
