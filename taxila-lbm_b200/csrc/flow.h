// flow.h -- host-side declarations shared by the translation units of libtaxila_gpu.so
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "stream_kernel.cuh"

namespace txg {

// One set of kernel entry points per (lattice, S, MRT, ISO) combination; the instantiations are
// spread over inst_*.cu so they compile in parallel.
struct KernelSet {
  // hot path: one lane per (fluid node, component); (first, count) select the positions
  void (*moments)(Grid, Phys, const double *, double *, long long, long long);
  void (*collide)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *,
                  const uint32_t *, const double *, long long, long long, int);
  // the same step fed by bulk asynchronous copies (stream_kernel.cuh): persistent blocks
  void (*collide_stream)(Grid, Phys, const double *, double *, const double *, const uint32_t *, const uint32_t *,
                         const uint32_t *, const double *, long long, long long, int, int);
  int stream_smem;  // dynamic shared memory per block of collide_stream
  int stream_threads, stream_pb;  // block size and positions per chunk of collide_stream
  void (*build_nbr)(Grid, uint32_t *);
  void (*halo_unpack)(Grid, double *, const double *, long long, int, const uint32_t *, long long, long long, int);
  // set-up and export
  void (*fi_init)(Grid, Phys, double *, const double *, const double *, const double *, const uint32_t *,
                  const uint32_t *, const uint8_t *, int, int);
  void (*export_state)(Grid, Phys, const double *, const double *, const uint32_t *, const uint32_t *,
                       const uint8_t *, double *, double *, double *, double *, double *, double *, double, int, int);
  void (*build_masks)(Grid, const uint8_t *, uint32_t *, uint32_t *, int *);
  void (*build_wallrec)(Grid, Phys, const uint8_t *, const uint32_t *, const uint32_t *, double *);
  int npw;       // fluid nodes per warp of the hot kernels (32 / S)
  int ff_words;  // u32 words of ffmask per node (0 for isotropy order 4)
  const char *name;
};

template <class L, int S, bool MRT, int ISO>
KernelSet make_kernel_set(const char *name) {
  KernelSet k;
  k.moments = k_moments<L, S>;
  k.collide = k_collide<L, S, MRT, ISO>;
  k.collide_stream = k_collide_stream<L, S, MRT, ISO>;
  k.stream_smem = (int)sizeof(StreamSmem<L, S>) + 128;
  k.stream_threads = STREAM_THREADS;
  k.stream_pb = StreamStage<L, S>::PB;
  k.halo_unpack = k_halo_unpack<L, S>;
  k.fi_init = k_fi_init<L, S, ISO>;
  k.export_state = k_export<L, S, ISO>;
  k.build_masks = k_build_masks<L, ISO>;
  k.build_wallrec = k_build_wallrec<L, S, ISO>;
  k.build_nbr = k_build_nbr<L>;
  k.npw = Lanes<S>::NPW;
  k.ff_words = ISO == 4 ? 0 : ff_words<L>(ISO);
  k.name = name;
  return k;
}

// defined in inst_<lattice>_s<S>.cu; returns false if the combination is not built
bool kernel_set_d3q19_s1(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s2(bool mrt, int iso, KernelSet *out);
bool kernel_set_d3q19_s3(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s1(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s2(bool mrt, int iso, KernelSet *out);
bool kernel_set_d2q9_s3(bool mrt, int iso, KernelSet *out);

}  // namespace txg
