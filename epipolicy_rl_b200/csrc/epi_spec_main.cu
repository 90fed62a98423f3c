// epi_spec_main.cu — translation unit of a scenario-specialised build of the cell kernel.
// Compiled by epipolicy_rl_b200/spec.py:  nvcc ... -DEPI_SPEC_HEADER='"<generated header>"' -shared
// The generated header (epi_spec_source, include/epi_b200.h) defines EPI_SPEC, EPI_SPEC_KEY and the
// scenario's per-cell program in namespace epi::spec; epi_cell.cuh then compiles its specialised path.
#include EPI_SPEC_HEADER
#if EPI_SPEC_BIG
#include "epi_env.cuh"
#define EPI_SPEC_KERNEL epi::env::env_kernel<epi::spec::real>
#if EPI_GRP_THREADS > 0
#include "epi_grp.cuh"
#endif
#else
#include "epi_cell.cuh"
#define EPI_SPEC_KERNEL epi::cell::cell_kernel<epi::spec::real>
#endif

extern "C" const char* epi_spec_key() { return EPI_SPEC_KEY; }

#ifdef EPI_HOST_EMU
// tests only: the same specialised code on the host (tests/emu/cuda_runtime.h), warps / CTAs as lock-stepped pthreads
extern "C" int epi_spec_launch(const DevPlan* P, const DevState* S, const epi::StepArgs* a, int grid, int threads,
                               size_t smem, void*) {
#if EPI_SPEC_BIG
  emu::run_warps(grid, threads, smem, [&](char* sm, int tid, int cta) {
    epi::env::Bx b = epi::env::bind(*P, *a, sm, tid, threads, cta);
    epi::env::step_cta<epi::spec::real>(*P, *S, *a, b, cta, grid);
  });
#else
  emu::run_warps(grid, threads, smem, [&](char* sm, int lane, int group) {
    epi::cell::step_warp<epi::spec::real>(*P, *S, *a, sm, lane, group);
  });
#endif
  return 0;
}
#else
extern "C" int epi_spec_launch(const DevPlan* P, const DevState* S, const epi::StepArgs* a, int grid, int threads,
                               size_t smem, void* stream) {
  auto kern = EPI_SPEC_KERNEL;
  // the attribute is a property of the function, set when the size first appears (per device): a launch that is being
  // captured into a CUDA graph makes no other API call than the launch itself
  static size_t set_smem[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || set_smem[dev] != smem + 1) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) set_smem[dev] = smem + 1;
  }
  kern<<<grid, threads, smem, (cudaStream_t)stream>>>(*P, *S, *a);
  return (int)cudaGetLastError();
}
#endif

#if EPI_SPEC_BIG && EPI_GRP_THREADS > 0
// the env-group kernel (epi_grp.cuh): `grid` = n_groups x S CTAs that must all be resident (they wait on one another)
#ifdef EPI_HOST_EMU
extern "C" int epi_spec_launch_grp(const DevPlan* P, const DevState* S, const epi::StepArgs* a, int grid, int threads,
                                   size_t smem, void*) {
  const int Sg = P->gl.S, ng = grid / Sg;
  emu::run_groups(ng, Sg, threads, smem, [&](char* sm, int tid, int block) {
    epi::grp::Gx b = epi::grp::bind(*P, *a, sm, tid, threads, block);
    epi::grp::step_group(*P, *S, *a, b, block / Sg, ng);
  });
  return 0;
}
#else
extern "C" int epi_spec_launch_grp(const DevPlan* P, const DevState* S, const epi::StepArgs* a, int grid, int threads,
                                   size_t smem, void* stream) {
  auto kern = epi::grp::grp_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  void* args[3] = {(void*)P, (void*)S, (void*)a};
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), args, smem, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaGetLastError();
}
#endif
#endif
