// flow.h -- host-side declarations shared by the translation units of libtaxila_gpu.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "band_kernel.cuh"
#include "stage_kernel.cuh"
#include "fused_kernel.cuh"

namespace txg {

// One set of kernel entry points per (lattice, S, MRT, ISO) combination; the instantiations are
// spread over inst_*.cu so they compile in parallel.
struct KernelSet {
  // hot path: one lane per (fluid node, component); (first, count) select the positions
  // (the *_pair forms cover two runs of positions in one launch: the bottom and top boundary planes of a z-slab)
  void (*moments)(Grid, Phys, const double *, double *, double *, long long, long long, long long, long long);
  void (*moments_pair)(Grid, Phys, const double *, double *, double *, long long, long long, long long, long long);
  void (*forces)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *, const uint32_t *,
                 const double *, double *, long long, long long, FaceDesc, const uint32_t *);
  void (*forces_face)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *, const uint32_t *,
                      const double *, double *, long long, long long, FaceDesc, const uint32_t *);
  void (*collide)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *, long long,
                  long long, const double *, FaceDesc, const uint32_t *);
  // the same over the fluid nodes of one box face (fused step with external face BCs)
  void (*collide_face)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *, long long,
                       long long, const double *, FaceDesc, const uint32_t *);
  // forces + collide in one kernel (order-4 stencil only; nullptr otherwise), fed by the full adjacency table
  void (*step_fused)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *,
                     const double *, long long, long long, int, long long, long long);
  void (*step_fused_pair)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *,
                          const double *, long long, long long, int, long long, long long);
  void (*fi_init_fused)(Grid, Phys, double *, const double *, const double *, const double *, const uint32_t *,
                        const uint32_t *, const double *, long long, long long);
  // one-pass step (opt-in, TXG_LAG=1): step_fused + the density sum of the next step in one launch (lag_schedule.h)
  void (*step_fused_lag)(Grid, Phys, LagMeta, const double *, double *, const double *, double *, const uint32_t *,
                         const uint32_t *, const double *, const LagRowDev *, unsigned *, unsigned *, const uint32_t *, int *, int);
  void (*step_fused_lag_tile)(Grid, Phys, LagMeta, const double *, double *, const double *, double *, const uint32_t *,
                              const uint32_t *, const double *, const LagRowDev *, unsigned *, unsigned *, const uint32_t *, int *,
                              int);  // + density tiles (TXG_LAG=1 TXG_RHOTILE=1)
  void (*build_rtab_lag)(Grid, const uint32_t *, int, const uint32_t *, long long, int, uint32_t *);
  int (*upload_lag_rows)(const void *rows, size_t bytes, cudaStream_t s);  // into this translation unit's c_lag_rows
  // step_fused with the stencil's neighbour densities staged in shared memory by bulk copies (opt-in, TXG_RHOTILE=1)
  void (*step_fused_tile)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *,
                          const double *, const uint32_t *, int *, long long, long long, int);
  void (*build_rtab)(Grid, const uint32_t *, int, long long, uint32_t *);
  // band blocks (band_kernel.cuh): step_fused without adjacency table, neighbour densities from shared-memory windows
  void (*step_band)(Grid, Phys, BandParams, const double *, double *, const double *, const double *, int, int, int, int *);
  // pull form (band_kernel.cuh): collided populations at their own nodes, gathered at the start of the next step
  void (*step_band_pull)(Grid, Phys, BandParams, const double *, double *, const double *, const double *, int, int, int, int *);
  void (*moments_pull)(Grid, Phys, BandParams, const double *, double *, double *, double *, long long, long long);
  void (*pull_stream)(Grid, Phys, BandParams, const double *, double *, double *, double *, long long, long long);
  int (*set_band_pull_smem)(int bytes);
  // tile-staged forces of the wide stencils (orders 8, 10): k_forces_tile (hot_kernels.cuh)
  void (*forces_tile)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *, const double *, double *, int, int,
                      int);
  // ... and the whole K2 of the wide stencils in one kernel (forces from the tile, then collide + push): k_step_tile
  void (*step_tile)(Grid, Phys, const double *, double *, const double *, const double *, const uint32_t *, const uint32_t *,
                    const uint32_t *, const double *, int, int, int);
  int (*set_forces_tile_smem)();
  int forces_tile_smem, forces_tile_tx, forces_tile_ty;
  // staged form (stage_kernel.cuh): k_step_fused with its streamed rows fetched by bulk copies into a double buffer
  // staged K2 in three block shapes (STAGE_VARIANTS: 4 warps x 4 blocks per SM, 6 x 2, 12 x 1; stage_kernel.cuh)
  void (*step_stage[3])(Grid, Phys, const CUtensorMap, const CUtensorMap, double *, const double *, const double *, long long, long long,
                        long long, int, int, const double *, const uint32_t *, const unsigned char *, const uint32_t *, int);
  // 4 warps x 4 blocks with the compressed adjacency records (AdjcGeom) instead of the adjacency + mask rows; their builder
  void (*step_stage_adjc)(Grid, Phys, const CUtensorMap, const CUtensorMap, double *, const double *, const double *, long long, long long,
                          long long, int, int, const double *, const uint32_t *, const unsigned char *, const uint32_t *, int);
  void (*export_diag_fused)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *, const double *, long long, long long,
                            double *, double *, double *);
  void (*build_adjc)(Grid, const uint32_t *, const uint32_t *, unsigned char *, long long, long long);
  int adjc_rec_bytes, stage_smem_adjc;
  // the same with blocks that take over the next block of the grid (cluster launch control)
  void (*step_stage_clc[3])(Grid, Phys, const CUtensorMap, const CUtensorMap, double *, const double *, const double *, long long, long long,
                            long long, int, int);
  int stage_warps[3], stage_smem_v[3], stage_blocks_v[3];
  int stage_item, stage_rows_f, stage_rows_a;  // box of the staged tensors: positions per item, population rows, adjacency + mask rows
  int (*set_stage_attrs)(int variant, int carveout_pct);  // dynamic shared memory size + carve-out of step_stage[variant] and its clc form
  int (*set_band_smem)(int bytes);  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) of step_band
  int band_threads, band_windows;   // block size; density windows per component (3 in 3-D, 1 in 2-D)
  int rtab_groups;  // window starts per block (RhoTile<L>::NG)
  void (*build_nbr_all)(Grid, uint32_t *);
  void (*build_nbr)(Grid, uint32_t *);
  void (*halo_unpack)(Grid, double *, const double *, long long, int, const uint32_t *, long long, long long, int);
  void (*halo_pack)(Grid, const double *, double *, long long, long long, int);
  // set-up and export
  void (*fi_init)(Grid, Phys, double *, const double *, const double *, const double *, const uint32_t *,
                  const uint32_t *, const uint8_t *, int, int);
  void (*export_state)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *,
                       const uint8_t *, const double *, const double *, int, double *, double *, double *, double *, double *, double *,
                       double, int, int);
  void (*build_masks)(Grid, const uint8_t *, uint32_t *, uint32_t *, int *);
  void (*build_wallrec)(Grid, Phys, const uint8_t *, const uint32_t *, const uint32_t *, double *);
  int fused_threads;  // block size of step_fused
  int npw;       // fluid nodes per warp of the hot kernels (32 / S)
  int ncen;      // rows of the adjacency table (centre directions)
  int ff_words;  // u32 words of ffmask per node (0 for isotropy order 4)
  const char *name;
};

template <class L, int S, bool MRT, int ISO>
KernelSet make_kernel_set(const char *name) {
  KernelSet k;
  k.moments = k_moments<L, S, false>;
  k.moments_pair = k_moments<L, S, true>;
  k.forces = k_forces<L, S, ISO, false>;
  k.forces_face = k_forces<L, S, ISO, true>;
  if constexpr (ISO != 4) {
    k.forces_tile = k_forces_tile<L, S, ISO>;
    k.step_tile = k_step_tile<L, S, MRT, ISO>;
    k.forces_tile_smem = S * ForceTile<L, ISO>::BOX * (int)sizeof(double);
    k.forces_tile_tx = ForceTile<L, ISO>::TX;
    k.forces_tile_ty = ForceTile<L, ISO>::TY;
    k.set_forces_tile_smem = []() -> int {
      cudaError_t e = cudaFuncSetAttribute(k_forces_tile<L, S, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           S * ForceTile<L, ISO>::BOX * (int)sizeof(double));
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_step_tile<L, S, MRT, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 S * ForceTile<L, ISO>::BOX * (int)sizeof(double));
      return (int)e;
    };
  } else {
    k.forces_tile = nullptr;
    k.step_tile = nullptr;
    k.set_forces_tile_smem = nullptr;
    k.forces_tile_smem = k.forces_tile_tx = k.forces_tile_ty = 0;
  }
  k.collide = k_collide<L, S, MRT, false>;
  k.collide_face = k_collide<L, S, MRT, true>;
  k.halo_unpack = k_halo_unpack<L, S>;
  k.halo_pack = k_halo_pack<L, S>;
  k.fi_init = k_fi_init<L, S, ISO>;
  k.export_state = k_export<L, S, ISO>;
  k.build_masks = k_build_masks<L, ISO>;
  k.build_wallrec = k_build_wallrec<L, S, ISO>;
  k.build_nbr = k_build_nbr<L>;
  k.build_nbr_all = k_build_nbr_all<L>;
  if constexpr (ISO == 4) {
    k.step_fused = k_step_fused<L, S, MRT, false>;
    k.step_fused_pair = k_step_fused<L, S, MRT, true>;
    k.fi_init_fused = k_fi_init_fused<L, S>;
    k.export_diag_fused = k_export_diag_fused<L, S>;
    k.step_fused_lag = k_step_fused_lag<L, S, MRT, false>;
    if constexpr (S <= 3) {  // (the density tiles of the opt-in experiments are static shared memory: S * 12 KB)
      k.step_fused_lag_tile = k_step_fused_lag<L, S, MRT, true>;
      k.build_rtab_lag = k_build_rtab_lag<L>;
      k.step_fused_tile = k_step_fused_tile<L, S, MRT>;
      k.build_rtab = k_build_rtab<L>;
      k.rtab_groups = RhoTile<L>::NG;
    } else {
      k.step_fused_lag_tile = nullptr;
      k.build_rtab_lag = nullptr;
      k.step_fused_tile = nullptr;
      k.build_rtab = nullptr;
      k.rtab_groups = 0;
    }
    k.step_band = k_step_band<L, S, MRT, false>;
    k.set_band_smem = [](int bytes) -> int {
      return (int)cudaFuncSetAttribute(k_step_band<L, S, MRT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    k.step_band_pull = k_step_band<L, S, MRT, true>;
    k.set_band_pull_smem = [](int bytes) -> int {
      return (int)cudaFuncSetAttribute(k_step_band<L, S, MRT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    k.step_stage[0] = k_step_stage<L, S, MRT, 4>;
    k.step_stage[1] = k_step_stage<L, S, MRT, 6>;
    k.step_stage[2] = k_step_stage<L, S, MRT, 12>;
    k.step_stage_adjc = k_step_stage<L, S, MRT, 4, true>;
    k.build_adjc = k_build_adjc<L, S>;
    k.step_stage_clc[0] = k_step_stage_clc<L, S, MRT, 4>;
    k.step_stage_clc[1] = k_step_stage_clc<L, S, MRT, 6>;
    k.step_stage_clc[2] = k_step_stage_clc<L, S, MRT, 12>;
    k.set_stage_attrs = [](int variant, int carveout_pct) -> int {
      auto set2 = [&](auto *kernel, auto *kernel_clc, int bytes) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel_clc, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel_clc, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
        return e;
      };
      if (variant == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_step_stage<L, S, MRT, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, StageGeom<L, S, 4, true>::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_stage<L, S, MRT, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
        if (e != cudaSuccess) return (int)e;
        return (int)set2(k_step_stage<L, S, MRT, 4>, k_step_stage_clc<L, S, MRT, 4>, StageGeom<L, S, 4>::SMEM_BYTES);
      }
      if (variant == 1) return (int)set2(k_step_stage<L, S, MRT, 6>, k_step_stage_clc<L, S, MRT, 6>, StageGeom<L, S, 6>::SMEM_BYTES);
      return (int)set2(k_step_stage<L, S, MRT, 12>, k_step_stage_clc<L, S, MRT, 12>, StageGeom<L, S, 12>::SMEM_BYTES);
    };
    k.moments_pull = k_moments_pull<L, S, false>;
    k.pull_stream = k_moments_pull<L, S, true>;
    k.upload_lag_rows = [](const void *rows, size_t bytes, cudaStream_t s) -> int {
      return (int)cudaMemcpyToSymbolAsync(c_lag_rows, rows, bytes, 0, cudaMemcpyDeviceToDevice, s);
    };
  } else {
    k.step_fused = nullptr;
    k.step_fused_pair = nullptr;
    k.fi_init_fused = nullptr;
    k.export_diag_fused = nullptr;
    k.step_fused_lag = nullptr;
    k.step_fused_lag_tile = nullptr;
    k.build_rtab_lag = nullptr;
    k.upload_lag_rows = nullptr;
    k.step_fused_tile = nullptr;
    k.build_rtab = nullptr;
    k.rtab_groups = 0;
    k.step_band = nullptr;
    k.set_band_smem = nullptr;
    k.step_band_pull = nullptr;
    k.set_band_pull_smem = nullptr;
    k.moments_pull = nullptr;
    k.pull_stream = nullptr;
    for (int v = 0; v < 3; ++v) k.step_stage[v] = nullptr, k.step_stage_clc[v] = nullptr;
    k.step_stage_adjc = nullptr;
    k.build_adjc = nullptr;
    k.set_stage_attrs = nullptr;
  }
  k.stage_warps[0] = 4, k.stage_warps[1] = 6, k.stage_warps[2] = 12;
  k.stage_smem_v[0] = StageGeom<L, S, 4>::SMEM_BYTES, k.stage_smem_v[1] = StageGeom<L, S, 6>::SMEM_BYTES, k.stage_smem_v[2] = StageGeom<L, S, 12>::SMEM_BYTES;
  k.stage_blocks_v[0] = StageGeom<L, S, 4>::BLOCKS_PER_SM, k.stage_blocks_v[1] = StageGeom<L, S, 6>::BLOCKS_PER_SM;
  k.stage_blocks_v[2] = StageGeom<L, S, 12>::BLOCKS_PER_SM;
  k.stage_item = StageGeom<L, S>::ITEM;
  k.adjc_rec_bytes = AdjcGeom<L, S>::REC_BYTES;
  k.stage_smem_adjc = StageGeom<L, S, 4, true>::SMEM_BYTES;
  k.stage_rows_f = StageGeom<L, S>::NF;
  k.stage_rows_a = StageGeom<L, S>::NA;
  k.fused_threads = TXG_FUSED_THREADS;
  k.band_threads = TXG_BAND_THREADS;
  k.band_windows = BandGeom<L>::NP;
  k.npw = Lanes<S>::NPW;
  k.ncen = num_centres<L>();
  k.ff_words = ISO == 4 ? 0 : ff_words<L>(ISO);
  k.name = name;
  return k;
}

// defined in inst_<lattice>_s<S>.cu; returns false if the combination is not built
bool kernel_set_d3q19_s1(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s2(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s3(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s4(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s5(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s1(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s2(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s3(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s4(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s5(bool mrt, int iso, KernelSet *out);

}  // namespace txg
