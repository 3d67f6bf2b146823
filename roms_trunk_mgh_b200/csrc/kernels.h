// Host-side launch wrappers of the sm_100a kernels (one per ROMS routine on the main3d path).
#pragma once
#include "dev.cuh"

namespace rb {
void launch_set_massflux(const Par& p, const Flds& f, cudaStream_t s);
void launch_rho_eos(const Par& p, const Flds& f, cudaStream_t s);
void launch_set_vbc(const Par& p, const Flds& f, cudaStream_t s);
void launch_omega(const Par& p, const Flds& f, cudaStream_t s);
void launch_wvelocity(const Par& p, const Flds& f, int Ninp, cudaStream_t s);
void launch_set_zeta(const Par& p, const Flds& f, cudaStream_t s);
void launch_set_depth(const Par& p, const Flds& f, cudaStream_t s);
void launch_ana_vmix(const Par& p, const Flds& f, cudaStream_t s);
void launch_bvf_mix(const Par& p, const Flds& f, cudaStream_t s);         // bvf_mix.F
void launch_bulk_flux(const Par& p, const Flds& f, cudaStream_t s);       // bulk_flux.F (whole tile only: not split-launch aware)
void launch_lmd_vmix(const Par& p, const Flds& f, cudaStream_t s);        // lmd_vmix.F + lmd_skpp.F (split-launch aware)
void launch_pre_step3d_t(const Par& p, const Flds& f, cudaStream_t s);    // tracer predictor (halo-exchanged: split-launch aware)
void launch_pre_step3d_uv(const Par& p, const Flds& f, cudaStream_t s);   // momentum loading
void launch_prsgrd(const Par& p, const Flds& f, int dj_gradps, cudaStream_t s);
void launch_t3dmix2_s(const Par& p, const Flds& f, cudaStream_t s);
void launch_t3dmix2_geo(const Par& p, const Flds& f, cudaStream_t s);
void launch_t3dmix4_s(const Par& p, const Flds& f, cudaStream_t s);       // t3dmix4_s.h (TS_DIF4)
void launch_rhs3d(const Par& p, const Flds& f, cudaStream_t s);
void launch_uv3dmix2(const Par& p, const Flds& f, cudaStream_t s);
// x != nullptr: this sub-step also pulls / pushes its xi-halo through NVLink peer memory (dev.cuh Xchg)
void launch_step2d(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x = nullptr);
// The whole LOOP_2D in one cooperative launch (k_step2d_loop.cu); false / 0: the tile is too large for it
bool launch_step2d_loop(const Par& p, const Flds& f, cudaStream_t s, const Xchg* x, const LoopCtl& ctl);
int step2d_loop_ctas(const Par& p);
// set_avg_tile (set_avg.F): mode 0 initialise, 1 accumulate; norm: scale by fac and fill the periodic images
void launch_set_avg(const Par& p, const Flds& f, int mode, int norm, double fac, int Kout, int Nout, cudaStream_t s);
void launch_step3d_uv(const Par& p, const Flds& f, cudaStream_t s);
void launch_step3d_t(const Par& p, const Flds& f, cudaStream_t s);
// diag: partial[] must hold 16 doubles per block row; out16 on device
void launch_diag(const Par& p, const Flds& f, double* partial, double* out16, int knew, cudaStream_t s);
int diag_partial_doubles(const Par& p);
// number of kernel launches each wrapper performs (for the gpu_launches claim)
}  // namespace rb
