// engine.h -- internal declarations shared by the engine translation units.
#pragma once
#include "frame.h"
#include "obs.h"
#include "rt.h"

namespace mcpm {

// Performance hint: the particle arrays given to the composite operators are a px x py x pz lattice in C order (z
// fastest, regular_pos of bricks.py:593-603), smoothly displaced.  Enables the brick-tiled scatter (brick.cu); never
// affects results (stray particles take the generic path).
struct Lattice {
  int px = 0, py = 0, pz = 0;
};

// Flat-sky redshift-space shift applied INSIDE a paint (bricks.py:781-792 in cell units): the particle is deposited at
// pos + (vel . los) * coef * los without that position ever being written to memory.  vel == NULL: no shift.
struct ObsShift {
  const float* vel = nullptr;
  float lx = 0.f, ly = 0.f, lz = 0.f, coef = 0.f;
  // the general observation transform (obs.h) instead: a HOST pointer, copied into the generic assignment kernels'
  // arguments (paint.cu); the brick-tiled paint does not implement it and is bypassed
  const ObsGen* gen = nullptr;
};

// Performance knobs (never change which result is computed).  mcpm_tune sets the process-wide defaults, which every
// engine copies at creation; mcpm_engine_tune changes one engine.  Kernels read the knobs of the engine whose entry point
// is executing on the calling host thread (a thread-local pointer held for the duration of the call), so engines on
// other devices or threads -- one replica per GPU in one process, SURVEY 8b -- never see each other's settings; the
// stateless entry points use the defaults.
struct Tune {
  int side_zero = 0;       // clear the next step's scatter meshes inside the gather kernels instead of memsets
  int gather_minb = 4;     // resident CTAs per SM the cic4.cu gathers are compiled for (4 | 5 | 6)
  int gather_blocked = 0;  // cic4.cu gathers: one CTA per 256 consecutive particles instead of a grid-stride loop
  int brick = 1;           // brick-tiled shared-memory scatters (brick.cu) where they apply; 0: generic global atomics
  int brick_stream = 44;   // brick.cu: persistent, bulk-copy staged 3-channel scatter; value = tile row stride in words (44 | 48), 0: off
  int brick_stream1 = 0;   // ... also for the 1-channel density paint (measured slower than the per-brick kernel)
  int gather_tma = 1;      // gathers with bulk-copy staged particle arrays (cic4_tma.cu) where they apply
  int gather_seg = 32;     // particles per bulk copy there (32 | 64 | 128)
  // Both measured on a B200 in round 2 and left OFF (profiles/r2_tune_*.txt): the fused (y,z) kernel runs 84 / 96 us per
  // 256^3 mesh (R2C / C2R) against 65 us for cuFFT's two passes, whose second pass reads the first one's output from L2;
  // the 2 x 4 patch mapping of the gathers costs 0.9 ms per evaluation more than one pencil per CTA.
  int yzfft = 0;           // the batched (y,z) transforms as one fused kernel (yzfft.cu) where supported; 0: cuFFT 2-D plans
  int gather_brick = 0;    // cic4_tma.cu: a CTA's 8 warps take a 2 x 4 patch of lattice rows instead of one pencil
};
Tune& default_tune();  // api.cu
const Tune& tune();    // the executing engine's knobs, else the defaults
struct TuneScope {
  const Tune* prev;
  explicit TuneScope(const Tune* t);
  ~TuneScope();
};
int tune_set(Tune& t, const char* key, int value);  // MCPM_OK or MCPM_EINVAL (message set)

struct Engine {
  static constexpr int kR = 7;  // real scratch meshes (6 Hessian / 3 force + 1 density)
  static constexpr int kC = 7;  // half-spectrum scratch meshes
  int nx, ny, nz, nzc;
  int device = 0;  // CUDA device the plans and scratch live on
  Lattice lat;     // optional particle-order hint
  Tune tune;       // this engine's performance knobs (copied from the process defaults at creation)
  int rel = 0;     // 1: the composite operators take lattice-relative positions (frame.h); needs `lat`
  // the frame the particle kernels are given: NULL (absolute) or the lattice `lat` spanning this engine's mesh
  const Frame* frame(Frame& f) const {
    if (!rel) return nullptr;
    f = make_rel_frame(lat.px, lat.py, lat.pz, nx, ny, nz);
    return &f;
  }
  int64_t N, Nc;
  float invN;
  FftPlans* fft = nullptr;
  struct SlabFft* fft2d = nullptr;  // batched 2-D (y,z) plans per x-plane, for the fused x-transform path (xfft.cu)
  int fused_fft = 0;                // 1: pm_forces and its VJP run 2-D cuFFT + one fused x-pass kernel
  float* rbuf = nullptr;
  cfloat* cbuf = nullptr;
  size_t scratch_bytes = 0;
  float* r(int i) const { return rbuf + (int64_t)i * N; }
  cfloat* c(int i) const { return cbuf + (int64_t)i * Nc; }
};

Engine* engine_create(int nx, int ny, int nz);
void engine_destroy(Engine*);

// paint.cu.  kb_kcut > 0 selects the Kaiser-Bessel window with that cutoff (nbody.py:280-290) instead of `rectangular`.
int paint(stream_t, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny, int nz,
          int order, const float* scale, float shift, float* mesh, int accumulate, float kb_kcut = 0.0f,
          const Frame* fr = nullptr, const ObsShift* obs = nullptr);
int read(stream_t, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz, int order,
         const float* scale, float shift, float* out, float kb_kcut = 0.0f, const Frame* fr = nullptr);
int read_grad(stream_t, const float* pos, const float* const* meshes, int nmesh, const float* cot, int ncot,
              float cscale, const float* gw, int64_t np, int nx, int ny, int nz, int order, const float* scale,
              float shift, float* grad, int accumulate, float kb_kcut = 0.0f, const Frame* fr = nullptr);
int paint3(stream_t, const float* pos, const float* A, float ca, const float* B, float cb, int64_t np, int nx,
           int ny, int nz, int order, float* mesh3, int accumulate, const Frame* fr = nullptr);
int paint_vjp(stream_t, const float* pos, const float* weights, float wscalar, const float* mbar, int64_t np, int nx,
              int ny, int nz, int order, const float* scale, float shift, float* posbar, float* wbar,
              int accumulate, float kb_kcut = 0.0f, const Frame* fr = nullptr, const ObsShift* obs = nullptr,
              float* velbar = nullptr, int nshift = 1, int64_t mesh_stride = 0);
int kick_drift(stream_t, const float* pos, const float* vel, const float* fmesh3, int64_t np, int nx, int ny, int nz,
               int order, float alpha, float beta, float drift, float* pos_out, float* vel_out, float* force_out,
               const Frame* fr = nullptr);
int read_sites3(stream_t, const float* planar3, int64_t n, float* out);
int paint_sites3(stream_t, const float* A, float ca, const float* B, float cb, int64_t n, float* mesh3, int accumulate);
int axpy3(stream_t, const float* a, const float* b, float s, int64_t n3, float* out);
int lpt_combine(stream_t, const float* pos, const float* f1, const float* f2, float d1, float d2, float dv2,
                int64_t np, float* dpos, float* vel, float* pos_out);
int dot_accum(stream_t, const float* a, const float* b, int64_t n, double scale, double* out);
int axpby(stream_t, const float* x, float a, const float* y, float b, float c, int64_t n, float* out);
// bias.cu
struct BiasCoef {
  float b1, b2, bs2, b3, bds2, bs3, bn2, bnpar, fbp, fbpd, fbpd2, fbps2, fbn2p;
};
int bias_spectra(stream_t, const cfloat* dk, int nx, int ny, int nz, float cx, float cy, float cz, const float* inv_transfer,
                 cfloat* out);
int bias_spectra_T(stream_t, const cfloat* outbar, int nx, int ny, int nz, float cx, float cy, float cz,
                   const float* inv_transfer, cfloat* dkbar, int accumulate);
int shear_invariants(stream_t, const float* s5, int64_t n, float* out2);
int shear_invariants_vjp(stream_t, const float* s5, const float* out2bar, int64_t n, float* s5bar);
int bias_moments(stream_t, const float* vals, int K, float gs, const float* garr, int64_t np, double* mom);
int bias_weights(stream_t, const float* vals, int K, float gs, const float* garr, BiasCoef c, const double* mom, int64_t np,
                 float* weights, float* dvel);
int bias_weights_vjp(stream_t, const float* vals, int K, float gs, const float* garr, BiasCoef c, const double* mom,
                     const float* wbar, const float* dvelbar, int64_t np, double* msum, float* valsbar, double* coefbar,
                     float* gbar_arr);
int absmax_strided(stream_t, const float* x, int64_t n, int stride, float* out);
int yz_gradients(stream_t, cfloat* buf3, int xl, int ny, int nz, int grad_fd, int transpose);
int obs_reduce_slots(stream_t st, double* parbar, int64_t row);  // paint.cu
int radial_tables(stream_t st, const float* pos, int64_t np, ObsGen g, int ntab, const float* tabs, float* out);
int radial_tables_vjp(stream_t st, const float* pos, int64_t np, ObsGen g, int ntab, const float* tabs, const float* outbar,
                      float* posbar, double* tabbar);
int rsd_shift(stream_t, const float* pos, const float* vel, float lx, float ly, float lz, float coef, int64_t np,
              float* pos_out);
int rsd_shift_vjp(stream_t, const float* posbar, float lx, float ly, float lz, float coef, int64_t np, float* velbar,
                  int accumulate);

// Local ky block of a slab-decomposed half spectrum [nx, ny_loc, nz/2+1]; ny_loc = 0 means the full mesh.
struct SlabK {
  int ny_loc = 0, y0 = 0;
};

// Distributed-FFT building blocks for slab decomposition over `parts` ranks (fft.cu): batched 2-D transforms over (y,z)
// of the local x-planes, and 1-D transforms along x of the local ky block.  All unnormalised.
struct SlabFft;
SlabFft* slabfft_create(int nx, int ny, int nz, int parts);
void slabfft_destroy(SlabFft*);
int slabfft_r2c_yz(SlabFft*, stream_t, const float* in, cfloat* out, int nb);   // [nb, xl, ny, nz] -> [nb, xl, ny, nzc]
int slabfft_c2r_yz(SlabFft*, stream_t, cfloat* in, float* out, int nb);         // inverse, may overwrite its input
int slabfft_c2c_x(SlabFft*, stream_t, cfloat* data, int nb, int inverse);       // in place on [nb][nx, kyl, nzc]

// fourier.cu  (`norm` multiplies the output; the engine passes 1/N ahead of a raw C2R)
int force_spectra(stream_t, const cfloat* dk, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk = SlabK());
int force_spectra_T(stream_t, const cfloat* in3, cfloat* out1, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float kcut, int deconv_order, int half_weights, int accumulate, float norm, SlabK sk = SlabK());
int hessian_spectra(stream_t, const cfloat* dk, cfloat* out6, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float norm, SlabK sk = SlabK());
int hessian_spectra_T(stream_t, const cfloat* in6, cfloat* out1, int nx, int ny, int nz, int lap_fd, int grad_fd,
                      int half_weights, int accumulate, float norm, SlabK sk = SlabK());
int lpt2_source(stream_t, const float* h6, float* d2, int64_t n);
int lpt2_source_vjp(stream_t, const float* h6, const float* d2bar, float* hbar6, int64_t n);
int deconv(stream_t, const cfloat* in, cfloat* out, int nx, int ny, int nz, int order, float kb_kcut = 0.0f);
int interlace_combine(stream_t, const cfloat* in_m, cfloat* out, int m, int nx, int ny, int nz, float scale,
                      int deconv_order, SlabK sk = SlabK());
int interlace_combine_T(stream_t, const cfloat* in, cfloat* out_m, int m, int nx, int ny, int nz, float scale,
                        int deconv_order, float norm, SlabK sk = SlabK(), int half_weights = 1);
int chreshape_crop_xz_slab(stream_t, const cfloat* in, int inx, int iny, int inz, int rows, int y0, const cfloat* nyq_plane,
                           const float* row_w, cfloat* out, int onx, int onz, float scale);
int chreshape_crop_xz_slab_T(stream_t, const cfloat* outbar, int onx, int onz, int rows, int y0, const float* row_w,
                             cfloat* inbar, int inx, int iny, int inz, cfloat* nyq_plane_bar, float scale);
int hermitian_project(stream_t, cfloat* data, int nx, int ny, int nz, int batch);
int half_weight_axpy(stream_t, const cfloat* in, cfloat* out, int64_t nc, int nz, float a, int inverse, int accumulate);
int scale_spectrum(stream_t, const cfloat* in, const float* t, cfloat* out, int64_t nc);
int scale_real(stream_t, const float* in, float s, float* out, int64_t n);
int chreshape(stream_t, const cfloat* in, int inx, int iny, int inz, cfloat* out, int onx, int ony, int onz);
int chreshape_T(stream_t, const cfloat* outbar, int onx, int ony, int onz, cfloat* inbar, int inx, int iny, int inz);
int spectrum_bins(stream_t, const cfloat* m0, const cfloat* m1, int nx, int ny, int nz, double bx, double by, double bz,
                  const double* kedges, int n_edges, int deconv0, int deconv1, double* out, int ell = 0, double lx = 0.0, double ly = 0.0,
                  double lz = 0.0);
int rg2cgh(stream_t, const float* mesh, cfloat* out, int nx, int ny, int nz, float scale, const float* transfer);
int rg2cgh_vjp(stream_t, const cfloat* outbar, float* meshbar, int nx, int ny, int nz, float scale, const float* transfer);
int cgh2rg(stream_t, const cfloat* meshk, float* mesh, int nx, int ny, int nz, float inv_scale);
int hermitian_weights(stream_t, const cfloat* in, cfloat* out, int nx, int ny, int nz, int mode);

// cic4.cu: CIC kernels on the float4-interleaved vector mesh
int interleave3(stream_t, const float* planar3, float* mesh4, int64_t n);
int deinterleave3(stream_t, const float* mesh4, float* planar3, int64_t n);
int kick_drift4(stream_t, const float* pos, const float* vel, const float* fmesh4, int64_t np, int nx, int ny, int nz,
                float alpha, float beta, float drift, float* pos_out, float* vel_out, float* zero = nullptr,
                int64_t nzero = 0, const Frame* fr = nullptr);
int paint3v4(stream_t, const float* pos, float* A, const float* B, float cb, int store, float scale, int64_t np, int nx,
             int ny, int nz, float* mesh4, const Frame* fr = nullptr);
int read_grad4v(stream_t, const float* pos, const float* fmesh4, const float* rhobar, float* cot, float cscale,
                int scale_cot, float alpha_tail, int64_t np, int nx, int ny, int nz, float* grad, int accumulate,
                float* zero = nullptr, int64_t nzero = 0, const Frame* fr = nullptr, float dnext = 0.0f);

// halo.cu: halo exchange of a slab-decomposed mesh as kernels over peer memory
int halo_reduce_peer(stream_t, float* own, const float* prev, const float* next, int H, int xl, int64_t plane, int nlead, int h);
int halo_gather_peer(stream_t, float* own, const float* prev, const float* next, int H, int xl, int64_t plane, int nlead, int h);
int halo_gather4_peer(stream_t, float* fm4_ext, const float* F_own, const float* F_prev, const float* F_next, int H, int xl,
                      int64_t plane, int h);

// yzfft.cu (CUDA build only): fused two-pass (y,z) R2C / C2R on square planes of side 64, 128, 256; unnormalised
bool yzfft_supported(int ny, int nz);
int yzfft_r2c(stream_t, const float* in, cfloat* out, int ny, int nz, int nplanes);
int yzfft_c2r(stream_t, const cfloat* in, float* out, int ny, int nz, int nplanes);

// cic4_tma.cu (CUDA build only): the same two gathers with bulk-copy staged particle arrays; 1 handled, 0 not applicable
int kick_drift4_tma(stream_t, const float* pos, const float* vel, const float* fmesh4, int64_t np, int nx, int ny, int nz,
                    float alpha, float beta, float drift, float* pos_out, float* vel_out, const Frame* fr);
int read_grad4v_tma(stream_t, const float* pos, const float* fmesh4, const float* rhobar, float* cot, float cscale,
                    int scale_cot, float alpha_tail, int64_t np, int nx, int ny, int nz, float* grad, int accumulate,
                    const Frame* fr, float dnext);

// xfft.cu (CUDA build only): the x-passes of rfftn / irfftn fused with the force kernel, on [nx, ny_loc, nz/2+1]
bool xfuse_supported(int nx);
int xfuse_force(stream_t, const cfloat* in, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd, float kcut,
                int deconv_order, float norm, SlabK sk = SlabK());
int xfuse_force_T(stream_t, const cfloat* in3, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                  float kcut, int deconv_order, float norm, SlabK sk = SlabK());
int xfuse_force_peer(stream_t, const cfloat* const* in_peers, cfloat* const* out_peers, int npeer, int transpose, int nx,
                     int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, float kcut, int deconv_order,
                     float norm);
int xfuse_peer_mode(stream_t, int mode, const cfloat* const* in_peers, cfloat* const* out_peers, cfloat* k_local, int npeer,
                    int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, int half_weights, int accumulate,
                    float norm);
int xfuse_force_k(stream_t, const cfloat* dk, cfloat* out3, int nx, int ny, int nz, int lap_fd, int grad_fd, float kcut,
                  int deconv_order, float norm, SlabK sk = SlabK());
int xfuse_hessian_k(stream_t, const cfloat* dk, cfloat* out6, int nx, int ny, int nz, int lap_fd, int grad_fd,
                    float norm, SlabK sk = SlabK());
int xfuse_force_tk(stream_t, const cfloat* in3, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                   float kcut, int deconv_order, int half_weights, int accumulate, float norm, SlabK sk = SlabK());
int xfuse_hessian_tk(stream_t, const cfloat* in6, cfloat* out, int nx, int ny, int nz, int lap_fd, int grad_fd,
                     int half_weights, int accumulate, float norm, SlabK sk = SlabK());

// brick.cu (CUDA build only): return 1 if handled, 0 if the generic path must be taken, < 0 on error
int brick_paint_cic(stream_t, const Lattice&, const float* pos, const float* weights, float wscalar, float shift,
                    int64_t np, int nx, int ny, int nz, float* mesh, const Frame* fr = nullptr, const ObsShift* obs = nullptr);
int brick_paint3_cic(stream_t, const Lattice&, const float* pos, const float* A, float s, int64_t np, int nx, int ny,
                     int nz, float* mesh3, const Frame* fr = nullptr);

// engine.cu
int pm_forces(Engine*, stream_t, const float* pos, int64_t np, int order, int paint_deconv, int lap_fd, int grad_fd,
              float kcut, float* fmesh3, float* forces, bool rho_prezeroed = false);
int pm_forces_vjp(Engine*, stream_t, const float* pos, const float* fbar, float cscale, const float* fmesh3,
                  int64_t np, int order, int paint_deconv, int lap_fd, int grad_fd, float kcut, float* posbar,
                  int accumulate);
int pm_forces_mesh(Engine*, stream_t, const float* pos, const cfloat* dk, int64_t np, int order, int lap_fd,
                   int grad_fd, float kcut, float* forces);
int pm_forces2(Engine*, stream_t, const float* pos, const cfloat* dk, int64_t np, int order, int lap_fd, int grad_fd,
               float* forces, float* h6_out);
int lpt(Engine*, stream_t, const cfloat* dk, const float* pos, int64_t np, int lpt_order, int read_order, int lap_fd,
        int grad_fd, float d1, float d2, float dv2, float* dpos, float* vel, float* f1, float* f2, float* h6);
int lpt_vjp(Engine*, stream_t, const float* pos, int64_t np, int lpt_order, int read_order, int lap_fd, int grad_fd,
            float d1, float d2, float dv2, const float* dposbar, const float* velbar, const float* f1,
            const float* f2, const float* h6, cfloat* dkbar, double* coefbar, int accumulate);
int nbody_steps(Engine*, stream_t, float* pos, float* vel, int64_t np, int n_steps, const float* alpha,
                const float* beta, const float* drift_pre, const float* drift_post, int order, int paint_deconv,
                int lap_fd, int grad_fd, float* xk, float* vk, float* fm);
int nbody_steps_vjp(Engine*, stream_t, float* posbar, float* velbar, int64_t np, int n_steps, const float* alpha,
                    const float* beta, const float* drift_pre, const float* drift_post, int order, int paint_deconv,
                    int lap_fd, int grad_fd, const float* xk, const float* vk, const float* fm, const float* v0,
                    double* coefbar);
int nufft(Engine*, stream_t, const float* pos, const float* weights, float wscalar, int64_t np, const float* scale,
          int paint_order, int interlace_order, int paint_deconv, cfloat* out_k, float kb_kcut = 0.0f, const ObsShift* obs = nullptr);
int nufft_vjp(Engine*, stream_t, const float* pos, const float* weights, float wscalar, int64_t np,
              const float* scale, int paint_order, int interlace_order, int paint_deconv, const cfloat* outbar_k,
              float* posbar, float* weightsbar, float kb_kcut = 0.0f,
              const ObsShift* obs = nullptr, float* velbar = nullptr);

}  // namespace mcpm
