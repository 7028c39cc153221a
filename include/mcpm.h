/*
 * mcpm.h  --  C ABI of the B200 particle-mesh engine (libmcpm.so).
 *
 * Drop-in boundary for the hot path of hsimonfroy/montecosmo: the Python callables of `montecosmo/nbody.py`
 * that `model.py:23-25` / `bricks.py:10` import.  The reference has no FFI layer of its own (it is pure JAX), so
 * each entry point names the reference callable (file:line) whose arithmetic it replaces; an XLA-FFI / ctypes stub
 * binding these symbols is shown in INTEGRATION.md.
 *
 * Conventions
 *   - All array pointers are DEVICE pointers owned by the caller; nothing is allocated or freed on the call path
 *     (engine scratch and cuFFT plans are created once, in mcpm_engine_create).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and CUDA-graph capturable.
 *   - Row-major C order.  pos / vel: float32 [np, 3], cell units, any real value (wrapping is the callee's job).
 *     Real meshes: float32 [nx, ny, nz].  Half spectra: interleaved complex64 [nx, ny, nz/2+1], forward transform
 *     unnormalised, inverse divides by nx*ny*nz (numpy "backward" norm, as jnp.fft).  nz must be even.
 *   - Return value: 0 on success, an MCPM_E* code otherwise; mcpm_last_error() gives the thread-local message.
 *   - Cotangents of complex arrays use the convention  zbar = dL/dRe(z) + i dL/dIm(z)  (torch).  JAX's convention
 *     is the complex conjugate of it; the XLA-FFI shim conjugates on the way out.
 */
#ifndef MCPM_H_
#define MCPM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCPM_VERSION 100 /* 0.1.0 */

enum {
  MCPM_OK = 0,
  MCPM_EINVAL = 1, /* bad argument (shape, order, null pointer)            */
  MCPM_ECUDA = 2,  /* CUDA runtime error                                   */
  MCPM_ECUFFT = 3, /* cuFFT error                                          */
  MCPM_ENOMEM = 4, /* engine scratch allocation failed                     */
  MCPM_EUNSUP = 5  /* valid in the reference but not implemented here      */
};

/* finite-difference order selectors of invlaplace_hat / gradient_hat (nbody.py:109-163): 0 means np.inf */
enum { MCPM_FD_INF = 0, MCPM_FD_2 = 2, MCPM_FD_4 = 4 };

typedef struct mcpm_engine mcpm_engine; /* opaque: cuFFT plans + scratch for one mesh shape on one device */

int mcpm_version(void);
const char* mcpm_last_error(void);
/* number of engine kernels (not cuFFT's) launched by this process so far; reset != 0 zeroes the counter */
long long mcpm_launch_count(int reset);

/* Process-wide performance knobs (never change which result is computed).  Keys: "gather_minb" = 4 | 5 | 6, the
 * resident CTAs per SM the readout kernels are compiled for; "gather_blocked" = 0 | 1, one CTA per 256 consecutive
 * particles instead of a grid-stride loop;
 * "side_zero" = 0 | 1, clear the next step's scatter meshes inside the readout kernels instead of memsets;
 * "brick" = 0 | 1, the brick-tiled shared-memory scatters (default 1; 0 = generic global-atomic kernels);
 * "gather_tma" = 0 | 1, the step-loop gathers with bulk-copy staged particle arrays (csrc/cic4_tma.cu; default 1);
 * "gather_seg" = 32 | 64 | 128, particles per bulk copy there (default 32); "gather_brick" = 0 | 1, a CTA's 8 warps take a
 * 2 x 4 patch of lattice rows instead of one z-pencil (default 0: measured slower); "yzfft" = 0 | 1, the batched (y,z)
 * transforms of the fused-FFT path as one kernel with both passes on-chip (csrc/yzfft.cu; square planes of side 64, 128,
 * 256) instead of cuFFT's two-kernel 2-D plans (default 0: measured 84 / 96 us against cuFFT's 65 us per 256^3 mesh). */
int mcpm_tune(const char* key, int value);
/* The same knobs for ONE engine.  mcpm_tune sets the process-wide defaults, which an engine copies when it is created and
 * which the stateless entry points use; an engine's own knobs apply to its composite operators only, so replicas on
 * several devices / threads in one process do not interfere. */
int mcpm_engine_tune(mcpm_engine* eng, const char* key, int value);

/* Engine for real mesh shape (nx, ny, nz) on the current device.  max_batch = largest number of meshes transformed
 * in one call (6 covers 2LPT).  Scratch = (2*max_batch+2) meshes + cuFFT work area, allocated here, once. */
int mcpm_engine_create(int nx, int ny, int nz, mcpm_engine** out);
int mcpm_engine_destroy(mcpm_engine* eng);
size_t mcpm_engine_scratch_bytes(const mcpm_engine* eng);
/* Performance hint, never changes which result is computed: the particle arrays passed to this engine's composite
 * operators are a px x py x pz lattice in C order (regular_pos, bricks.py:593-603), smoothly displaced.  Enables the
 * brick-tiled shared-memory scatter; particles that strayed from their brick's tile take the generic path.
 * px = 0 clears the hint. */
int mcpm_engine_set_lattice(mcpm_engine* eng, int px, int py, int pz);
/* Lattice-relative positions (needs mcpm_engine_set_lattice).  With on != 0 every `pos` array this engine's composite
 * operators take or return (mcpm_pm_forces*, mcpm_lpt*, mcpm_nbody_steps*, mcpm_nufft* and the xk tape) holds the
 * DISPLACEMENT of particle p from its lattice site  site_a(p) = i_a * n_a / p_a  (regular_pos, bricks.py:593-603;
 * (i_x, i_y, i_z) = unravel(p, (px, py, pz))) instead of the absolute position: x = site + pos.  The reference adds
 * lpt's dpos to the sites once (nbody.py:984-985) and carries the float64 sum; in float32 the sum resolves 1.5e-5 cell
 * at x ~ 256 (6e-5 at 1024), which flips particles across cell faces where the CIC derivative jumps.  The kernels
 * combine the site (exact integers) with the small displacement instead, so the CIC fraction resolves ~1e-7 cell at
 * any mesh size.  Cotangents are unchanged (d/dx = d/dpos).  `pos` may be NULL where it is read-only: the particles sit
 * on their sites (lpt's reads at q).  on = 0 restores absolute positions. */
int mcpm_engine_set_relative(mcpm_engine* eng, int on);
/* Fused x-transform path of pm_forces and its VJP (2-D cuFFT per x-plane + one kernel doing the x-FFT, the force
 * kernel of nbody.py:591-603 and the inverse x-FFTs).  On by default where supported (nx in {64, 128, 256, 512, 1024});
 * on = 0 selects the 3-D cuFFT + separate multiply path, on = 1 returns MCPM_EUNSUP where unsupported.
 * Both paths compute the same operator (float32 rounding differs). */
int mcpm_engine_set_fused_fft(mcpm_engine* eng, int on);

/* Brick-tiled CIC scatters for lattice-ordered particles (need a matching mcpm_engine_set_lattice; else MCPM_EUNSUP).
 * These are what mcpm_pm_forces / mcpm_nbody_steps_vjp run under the hint; exposed for timing and tests.  Meshes are
 * accumulated into (not zeroed).  paint_lattice: mesh += w_p * wscalar * W (CIC).  paint3_lattice: the reverse-step
 * scatter, vbar += xbar * drift (stored, when xbar != NULL), mesh3[c] += scale * vbar[., c] * W into 3 planar meshes. */
/* The same two kernels without an engine, for a lattice that does not span the mesh: px x py x pz particles in C order
 * (spacing one cell), painted into an nx x ny x nz mesh with px <= nx, ... -- a slab-decomposed rank's particles and
 * its halo-extended mesh.  MCPM_EUNSUP when the geometry is not supported (mesh smaller than a tile, nz % 4 != 0). */
int mcpm_paint_brick(void* stream, int px, int py, int pz, const float* pos, const float* weights, float wscalar,
                     float shift, int64_t np, int nx, int ny, int nz, float* mesh);
int mcpm_paint3_brick(void* stream, int px, int py, int pz, const float* pos, float* vbar, const float* xbar,
                      float drift, float scale, int64_t np, int nx, int ny, int nz, float* mesh3);
int mcpm_paint_lattice(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                       int64_t np, float* mesh);
int mcpm_paint3_lattice(mcpm_engine* eng, void* stream, const float* pos, float* vbar, const float* xbar, float drift,
                        float scale, int64_t np, float* mesh3);

/* ---- mass assignment ------------------------------------------------------------------------------------------
 * Positions are transformed in-kernel as x' = x * scale[d] + shift before assignment (nufft's final->paint units,
 * nbody.py:569, and interlace's shift, nbody.py:524); pass scale = {1,1,1}, shift = 0 for plain paint / read.
 * order: 1 NGP, 2 CIC, 3 TSC, 4 PCS (`rectangular`, nbody.py:220-246); kernel_type 'kaiser_bessel': the *_kb entry
 * points below. */

/* paint (nbody.py:365-396): mesh[(id0+s) mod n] += w_p * prod_d W(id0_d + s_d - x'_d).
 * weights may be NULL (then every particle carries `wscalar`), else w_p = weights[p] * wscalar.
 * accumulate = 0 zeroes `mesh` first. */
int mcpm_paint(void* stream, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny,
               int nz, int order, const float scale[3], float shift, float* mesh, int accumulate);

/* read (nbody.py:398-427): out[p] = sum_s mesh[(id0+s) mod n] * prod_d W(...);  nmesh meshes at once:
 * mesh = [nmesh, nx, ny, nz] planes, out = [np, nmesh] interleaved (nmesh = 3 gives a force array [np, 3]);
 * nmesh in 1..4, or 7 | 9 (the field sets of mcpm_bias_weights). */
int mcpm_read(void* stream, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz,
              int order, const float scale[3], float shift, float* out);

/* d/dx' of read: grad[p, a] = sum_m cot[p, m] * sum_s mesh_m[...] * dW_a * prod_{d != a} W_d, times scale[a].
 * cot may be NULL (all ones).  This one gather is the position-VJP of both read (cot = out_bar) and paint
 * (mesh = mesh_bar, cot = weights).  accumulate = 1 adds into grad. */
int mcpm_read_grad(void* stream, const float* pos, const float* mesh, int nmesh, const float* cot, int64_t np,
                   int nx, int ny, int nz, int order, const float scale[3], float shift, float* grad,
                   int accumulate);

/* VJP of paint from the mesh cotangent, one gather: weightsbar[p] = wscalar * read(mesh_bar)[p] and
 * posbar[p,a] = w_p * scale[a] * sum mesh_bar * dW_a * prod_{d != a} W_d (each output nullable). */
int mcpm_paint_vjp(void* stream, const float* pos, const float* weights, float wscalar, const float* mesh_bar,
                   int64_t np, int nx, int ny, int nz, int order, const float scale[3], float shift, float* posbar,
                   float* weightsbar, int accumulate);

/* ---- Kaiser-Bessel window (kernel_type = 'kaiser_bessel': nbody.py:280-290 window, 293-312 transform, selected at
 * nbody.py:321-322, 383-384, 415-416).  The same five operators with
 *   W(s) = I0(kc sqrt(1 - (2 s / order)^2)) / (order sinh(kc) / kc),  kc = kcut * order / 2,
 * on the same neighbours as `rectangular` of that order (nbody.py:375-376); kcut = optim_kcut(oversamp) =
 * 0.98 pi (2 - 1/oversamp) (nbody.py:357-363) is computed by the caller.  The position derivative is finite at the
 * edge of the support (s = order/2), where autodiff of the reference's sqrt would return inf. */
int mcpm_paint_kb(void* stream, const float* pos, const float* weights, float wscalar, int64_t np, int nx, int ny,
                  int nz, int order, float kcut, const float scale[3], float shift, float* mesh, int accumulate);
int mcpm_read_kb(void* stream, const float* pos, const float* mesh, int nmesh, int64_t np, int nx, int ny, int nz,
                 int order, float kcut, const float scale[3], float shift, float* out);
int mcpm_read_grad_kb(void* stream, const float* pos, const float* mesh, int nmesh, const float* cot, int64_t np,
                      int nx, int ny, int nz, int order, float kcut, const float scale[3], float shift, float* grad,
                      int accumulate);
int mcpm_paint_vjp_kb(void* stream, const float* pos, const float* weights, float wscalar, const float* mesh_bar,
                      int64_t np, int nx, int ny, int nz, int order, float kcut, const float scale[3], float shift,
                      float* posbar, float* weightsbar, int accumulate);
/* deconv_paint with the Kaiser-Bessel transform (nbody.py:293-312, 321-322): out = in / prod_d kb_hat(k_d). */
int mcpm_deconv_kb(void* stream, const void* in, void* out, int nx, int ny, int nz, int order, float kcut);

/* three-channel paint (VJP of a 3-mesh read w.r.t. the meshes): mesh3[m] += vscale * vals3[p,m] * W, m = 0..2 */
int mcpm_paint3(void* stream, const float* pos, const float* vals3, float vscale, int64_t np, int nx, int ny, int nz,
                int order, float* mesh3, int accumulate);

/* ---- FFT (jnp.fft.rfftn / irfftn call sites nbody.py:589,603,620,627,630) -------------------------------------
 * batch meshes, contiguous planes.  mcpm_irfftn may overwrite its input (cuFFT C2R).  It accepts ANY complex input, as
 * jnp.fft.irfftn does (which returns the real part of the full inverse transform): the input is first projected onto
 * Hermitian symmetry on the planes kz = 0 / Nyquist (mcpm_hermitian_project), because cuFFT's C2R on inconsistent input
 * is algorithm dependent. */
int mcpm_rfftn(mcpm_engine* eng, void* stream, const float* in, void* out_c64, int batch);
int mcpm_irfftn(mcpm_engine* eng, void* stream, void* in_c64, float* out, int batch);

/* ---- fused Fourier-space passes -------------------------------------------------------------------------------
 * mcpm_force_spectra: out_j = -gradient_hat_j * invlaplace_hat * [gaussian_hat(kcut)] * [1/rectangular_hat^2]
 *                             * delta_k, j = 0..2   (nbody.py:591-603).  kcut <= 0 means inf; deconv_order 0 = none.
 * conj != 0 applies the complex conjugate of the kernel and SUMS the three inputs into one output
 * (the transpose pass of the backward force: in = [3] spectra, out = [1] spectrum). */
int mcpm_force_spectra(void* stream, const void* delta_k, void* out3, int nx, int ny, int nz, int lap_fd,
                       int grad_fd, float kcut, int deconv_order);
int mcpm_force_spectra_T(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int lap_fd,
                         int grad_fd, float kcut, int deconv_order, int half_weights, int accumulate);

/* mcpm_hessian_spectra: out_ij = gradient_hat_i * gradient_hat_j * invlaplace_hat * delta_k for
 * (i,j) in (00, 11, 22, 01, 02, 12)  (nbody.py:611-627).  _T: sum_ij conj(kernel_ij) * in_ij -> one spectrum. */
int mcpm_hessian_spectra(void* stream, const void* delta_k, void* out6, int nx, int ny, int nz, int lap_fd,
                         int grad_fd);
int mcpm_hessian_spectra_T(void* stream, const void* in6, void* out1, int nx, int ny, int nz, int lap_fd,
                           int grad_fd, int half_weights, int accumulate);

/* 2LPT source from the six real Hessian meshes (nbody.py:615-627): d2 = sum_{i<j} (h_ii h_jj - h_ij^2).
 * _vjp: hbar6 from d2bar and h6. */
int mcpm_lpt2_source(void* stream, const float* h6, float* d2, int64_t n);
int mcpm_lpt2_source_vjp(void* stream, const float* h6, const float* d2bar, float* hbar6, int64_t n);

/* deconv_paint on a half spectrum (nbody.py:315-334): out = in / prod_d sinc(k_d / 2pi)^order. */
int mcpm_deconv(void* stream, const void* in, void* out, int nx, int ny, int nz, int order);

/* interlace combine (nbody.py:523-526) fused with nufft's Jacobian and deconvolution (nbody.py:571-574):
 * out = scale / prod sinc^deconv_order * (1/m) sum_{i<m} in_i * exp(+i (i/m) (kx+ky+kz)).
 * _T (transpose, conj kernel): out_i = conj(...) * in for each i, then ready for N*C2R(./w'). */
int mcpm_interlace_combine(void* stream, const void* in_m, void* out, int m, int nx, int ny, int nz, float scale,
                           int deconv_order);
int mcpm_interlace_combine_T(void* stream, const void* in, void* out_m, int m, int nx, int ny, int nz, float scale,
                             int deconv_order);

/* ---- slab-decomposed building blocks (multi-GPU) ----------------------------------------------------
 * The mesh is split along x over `parts` ranks (xl = nx/parts local planes); k-space is kept split along ky (kyl =
 * ny/parts rows, layout [nx, kyl, nz/2+1]).  A distributed rfftn is  r2c_yz -> all-to-all (x-planes for ky rows, by the
 * caller over NCCL) -> c2c_x;  irfftn runs backwards.  All transforms unnormalised.  The *_slab Fourier passes are the
 * single-GPU ones evaluated on the local ky block [y0, y0 + ny_loc); `norm` multiplies the output (fold 1/N here). */
typedef struct mcpm_slabfft mcpm_slabfft;
int mcpm_slabfft_create(int nx, int ny, int nz, int parts, mcpm_slabfft** out);
int mcpm_slabfft_destroy(mcpm_slabfft* h);
int mcpm_slabfft_r2c_yz(mcpm_slabfft* h, void* stream, const float* in, void* out_c64, int nb);
int mcpm_slabfft_c2r_yz(mcpm_slabfft* h, void* stream, void* in_c64, float* out, int nb);
int mcpm_slabfft_c2c_x(mcpm_slabfft* h, void* stream, void* data_c64, int nb, int inverse);
int mcpm_force_spectra_slab(void* stream, const void* delta_k, void* out3, int nx, int ny, int nz, int ny_loc, int y0,
                            int lap_fd, int grad_fd, float kcut, int deconv_order, float norm);
int mcpm_force_spectra_T_slab(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                              int lap_fd, int grad_fd, float kcut, int deconv_order, int half_weights, int accumulate,
                              float norm);
int mcpm_hessian_spectra_slab(void* stream, const void* delta_k, void* out6, int nx, int ny, int nz, int ny_loc, int y0,
                              int lap_fd, int grad_fd, float norm);
int mcpm_hessian_spectra_T_slab(void* stream, const void* in6, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                                int lap_fd, int grad_fd, int half_weights, int accumulate, float norm);

/* interlace combine and its transpose on the local ky block (the slab-decomposed final paint of the model).  _T:
 * out_i = norm * conj(kernel_i) * in, divided by the Hermitian weight w' when half_weights != 0. */
int mcpm_interlace_combine_slab(void* stream, const void* in_m, void* out, int m, int nx, int ny, int nz, int ny_loc,
                                int y0, float scale, int deconv_order);
int mcpm_interlace_combine_T_slab(void* stream, const void* in, void* out_m, int m, int nx, int ny, int nz, int ny_loc,
                                  int y0, float scale, int deconv_order, int half_weights, float norm);
/* chreshape (utils.py:975-1013) as a crop on a spectrum split along ky (the slab-decomposed final paint with a finer paint
 * mesh, BASELINE C5): `in` = [inx, rows, inz/2+1] holds global ky rows y0 .. y0+rows-1 of an iny-row spectrum; x and kz
 * are cropped to onx <= inx, onz <= inz and multiplied by `scale` (and by row_w[jr], nullable); the rows themselves are
 * redistributed by the caller (a selection, plus one 1/sqrt2-weighted sum for the new ky Nyquist row).  Cropping kz
 * folds conj in[-kx, -ky, onz/2] into the new Nyquist plane: `nyq_plane` = that plane of the WHOLE input spectrum,
 * [inx, iny], gathered by the caller (NULL when onz == inz).  _vjp: the transpose in the real inner product; inbar is
 * zeroed here, the mirrored contributions are accumulated into nyq_plane_bar [inx, iny] for the caller to sum over
 * ranks and add to plane onz/2 of its rows. */
int mcpm_chreshape_crop_xz_slab(void* stream, const void* in, int inx, int iny, int inz, int rows, int y0,
                                const void* nyq_plane, const float* row_w, void* out, int onx, int onz, float scale);
int mcpm_chreshape_crop_xz_slab_vjp(void* stream, const void* outbar, int onx, int onz, int rows, int y0,
                                    const float* row_w, void* inbar, int inx, int iny, int inz, void* nyq_plane_bar,
                                    float scale);
/* In-place Hermitian projection of `batch` half spectra on the self-conjugate planes kz = 0 and kz = Nyquist:
 * A(i,j,l) <- (A(i,j,l) + conj A(-i,-j,l)) / 2 -- what jnp.fft.irfftn applies implicitly to its input. */
int mcpm_hermitian_project(void* stream, void* data_c64, int nx, int ny, int nz, int batch);
/* out (+)= a * w' * in  (inverse = 0)  or  a / w' * in  (inverse != 0) on any block of a half spectrum whose fastest
 * axis is the whole kz axis: the weights of mcpm_hermitian_weights for callers that keep their own 1/N. */
int mcpm_half_weight_axpy(void* stream, const void* in, void* out, int64_t nc, int nz, float a, int inverse,
                          int accumulate);

/* Halo exchange of a slab-decomposed mesh over peer memory (halo-plane summation after painting, halo fetch
 * before readout).  Rank r owns planes [halo, halo + xl) of its extended mesh [nlead][xl + 2 halo][plane] (plane = floats
 * per x-plane, channels included; nlead = 3 for three planar meshes).  prev_ext / next_ext are the SAME buffer of ranks
 * r-1 / r+1 mapped into this process (CUDA IPC / symmetric memory; own pointers when there is one rank: the periodic
 * wrap).  One kernel per exchange reads the neighbours' planes over NVLink where they lie: no packing, no collective.
 *   reduce : own[halo + j] += prev[halo + xl + j],  own[xl + j] += next[j]             j < halo   (after a paint)
 *   gather : own[j] = prev[xl + j],  own[halo + xl + j] = next[halo + j]               j < halo   (before a readout)
 *   gather4: builds the float4 force mesh {Fx, Fy, Fz, 0} [xl + 2 halo][plane cells] from the three planar meshes of
 *            OWNED planes [3][xl][plane cells] of this rank and its two neighbours (interleave + halo fetch in one pass).
 * The caller orders the ranks with a barrier before (the neighbours' planes are complete) and before reuse.
 * active (0 = halo): only the `active` halo planes next to the owned region are exchanged -- what a step needs while its
 * particles stay within active - 1 planes of their sites (the halo is sized for a = 1; early steps need a few planes):
 *   reduce : own[halo + j] += prev[halo + xl + j],  own[xl + halo - active + j] += next[halo - active + j]    j < active
 *   gather : own[halo - active + j] = prev[xl + halo - active + j],  own[halo + xl + j] = next[halo + j]      j < active */
int mcpm_halo_reduce_peer(void* stream, float* own_ext, const float* prev_ext, const float* next_ext, int halo, int xl,
                          int64_t plane, int nlead, int active);
int mcpm_halo_gather_peer(void* stream, float* own_ext, const float* prev_ext, const float* next_ext, int halo, int xl,
                          int64_t plane, int nlead, int active);
int mcpm_halo_gather4_peer(void* stream, float* fmesh4_ext, const float* f3_own, const float* f3_prev, const float* f3_next,
                           int halo, int xl, int64_t plane, int active);

/* Fused x-transform passes (CUDA build, nx in {64, 128, 256, 512, 1024}; MCPM_EUNSUP otherwise), on a half spectrum
 * [nx, ny_loc, nz/2+1] that has been transformed along (y,z) only:
 *   xfuse_force   : out3[j] = IFFT_x( force kernel_j * FFT_x(in) )            = mcpm_force_spectra between the x-passes
 *   xfuse_force_T : out1    = IFFT_x( sum_j conj(kernel_j) * FFT_x(in3[j]) )  = mcpm_force_spectra_T (no half weights)
 * Unnormalised transforms; `norm` multiplies the output.  The engine's pm_forces / reverse step run these between
 * batched 2-D cuFFT plans; with ny_loc < ny they serve a slab-decomposed mesh after the all-to-all. */
int mcpm_xfuse_supported(int nx); /* 1 if the fused x-transform exists for this nx in this build, else 0 */
int mcpm_xfuse_force_slab(void* stream, const void* in, void* out3, int nx, int ny, int nz, int ny_loc, int y0,
                          int lap_fd, int grad_fd, float kcut, int deconv_order, float norm);
/* The same two operators with the distributed transpose inside the kernel (slab decomposition over npeer <= 8 GPUs of
 * one NVLink domain): in_peers[r] / out_peers[r] are rank r's buffers [ncomp][nx/npeer][ny][nz/2+1] -- the output of its
 * local 2-D R2C / the input of its local 2-D C2R -- mapped into this process (CUDA IPC / symmetric memory).  This rank
 * transforms the columns of ky rows y0 .. y0+ny_loc-1, loading every x-plane from and storing it to its owner over
 * NVLink: no all-to-all, no pack / unpack passes.  transpose = 0: 1 -> 3 components, 1: 3 -> 1; 2: 1 -> 2 components
 * (F^_x, Phi^) and 3: 2 -> 1 (C^_x, g_y C^_y + g_z C^_z) -- the two-field forms whose y / z parts mcpm_yz_gradients applies
 * on the local planes, so that one field in four stays off the wire (plain kernel only: no kcut / deconvolution).  The
 * caller orders the ranks with a barrier before (inputs complete everywhere) and after (outputs landed everywhere). */
int mcpm_xfuse_force_peer(void* stream, const void* const* in_peers, void* const* out_peers, int npeer, int transpose,
                          int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, float kcut,
                          int deconv_order, float norm);
/* The general form: any of the six operators of the fused x-transform with its x-space side in the peers' memory.
 * mode 0 FORCE (1 -> 3), 1 FORCE_T (3 -> 1): both sides x-space, as mcpm_xfuse_force_peer.  2 FORCE_K (spectrum -> 3
 * x-space components), 3 HESS_K (spectrum -> 6: the 2LPT Hessian set 00 11 22 01 02 12): k_local = this rank's ky block
 * of the 3-D spectrum [nx, ny_loc, nz/2+1] is read, results go to out_peers (inputs of the owners' 2-D C2R).
 * 4 FORCE_TK (3 -> spectrum), 5 HESS_TK (6 -> spectrum): x-space planes are read from in_peers (outputs of the owners'
 * 2-D R2C), the cotangent spectrum is written (accumulate != 0: added) to k_local, with the Hermitian weights w'/N when
 * half_weights != 0.  Peer buffers are [ncomp][nx/npeer][ny][nz/2+1].  These make the slab-decomposed lpt / lpt_vjp
 * (nbody.py:634-667) free of all-to-alls, pack and unpack passes: the kernel is the transpose. */
int mcpm_xfuse_peer(void* stream, int mode, const void* const* in_peers, void* const* out_peers, void* k_local, int npeer,
                    int nx, int ny, int nz, int ny_loc, int y0, int lap_fd, int grad_fd, int half_weights, int accumulate,
                    float norm);
int mcpm_xfuse_force_T_slab(void* stream, const void* in3, void* out1, int nx, int ny, int nz, int ny_loc, int y0,
                            int lap_fd, int grad_fd, float kcut, int deconv_order, float norm);

/* chreshape (utils.py:975-1013): Hermitian- and mean-preserving Fourier crop / pad between real shapes. */
int mcpm_chreshape(void* stream, const void* in, int inx, int iny, int inz, void* out, int onx, int ony, int onz);

/* VJP of chreshape (transpose in the real inner product): outbar at the OUTPUT shape -> inbar at the INPUT shape. */
int mcpm_chreshape_vjp(void* stream, const void* outbar, int onx, int ony, int onz, void* inbar, int inx, int iny,
                       int inz);

/* Binned auto / cross power spectrum, monopole (metrics.py:121-182; multipoles: the _ell variant below).  For every half-spectrum element:
 * bin = np.digitize(|k|, kedges), k = 2 pi f / box_size; out[0][bin] += w', out[1][bin] += w' |k|,
 * out[2..3][bin] += w' Re / Im (m0 conj m1) after dividing m_i by rectangular_hat^deconv_i; w' = 1 on kz = 0 / Nyquist,
 * else 2.  m1 NULL = auto spectrum.  kedges: n_edges float64 on the device, increasing; out: [4][n_edges + 1] float64 on
 * the device, accumulated into.  The caller normalises (metrics.py:176-177). */
int mcpm_spectrum_bins(void* stream, const void* m0_c64, const void* m1_c64, int nx, int ny, int nz, double box_x,
                       double box_y, double box_z, const double* kedges, int n_edges, int deconv0, int deconv1,
                       double* out);

/* The same with a multipole (metrics.py:165-166): the power sums out[2..3] carry (2 ell + 1) L_ell(mu), mu = k . los /
 * |k| (0 at k = 0) for the unit line of sight los[3] (host; box_center / |box_center|, zero for a centred box);
 * out[0..1] (count, sum |k|) are unchanged.  ell = 0 equals mcpm_spectrum_bins. */
int mcpm_spectrum_bins_ell(void* stream, const void* m0_c64, const void* m1_c64, int nx, int ny, int nz, double box_x,
                           double box_y, double box_z, const double* kedges, int n_edges, int deconv0, int deconv1,
                           int ell, const double los[3], double* out);

/* rg2cgh / cgh2rg (utils.py:785-921): real Gaussian mesh [nx,ny,nz] (all sides even) <-> complex
 * Gaussian Hermitian half spectrum by permutation and reweighting; rg2cgh(N(0,I)) is distributed as rfftn(N(0,I)).
 * out = scale * [transfer *] P(mesh), scale = sqrt(N/2) for norm "backward", 1/sqrt(2) "ortho", 1/sqrt(2N) "forward";
 * `transfer` (nullable, real [nx,ny,nz/2+1]) fuses the sqrt-power multiply of samp2base_mesh (bricks.py:305-309).
 * _vjp: cotangent of the real mesh from the complex cotangent dL/dRe + i dL/dIm.  cgh2rg: mesh = inv_scale * P^-1. */
int mcpm_rg2cgh(void* stream, const float* mesh, void* out_c64, int nx, int ny, int nz, float scale,
                const float* transfer);
int mcpm_rg2cgh_vjp(void* stream, const void* outbar_c64, float* meshbar, int nx, int ny, int nz, float scale,
                    const float* transfer);
int mcpm_cgh2rg(void* stream, const void* meshk_c64, float* mesh, int nx, int ny, int nz, float inv_scale);

/* Hermitian weights w' (1 on kz = 0 / Nyquist planes, else 2) that turn rfftn / irfftn into each other's transpose:
 * mode 0: out = in * N / w'  (VJP of rfftn: xbar = irfftn(out));  mode 1: out = in * w' / N  (VJP of irfftn: ybar = out,
 * with in = rfftn(xbar)). */
int mcpm_hermitian_weights(void* stream, const void* in, void* out, int nx, int ny, int nz, int mode);

/* ---- observation-chain glue around the path ----------------------------------------------
 * out = a*x + b*y + c (y nullable);  *out_f64 += sum a[i]*b[i] (float64 accumulation on device);
 * flat-sky RSD in cell units (bricks.py:781-792): pos_out = pos + (vel . los) * coef * los, and its VJP w.r.t. vel. */
int mcpm_axpby(void* stream, const float* x, float a, const float* y, float b, float c, int64_t n, float* out);
int mcpm_dot(void* stream, const float* a, const float* b, int64_t n, double* out_f64);
/* The y and z force components from the potential on spectra already in x-space (a slab-decomposed caller then moves two
 * fields through the distributed x-transform instead of three): buf3 = [3][xl][ny][nz/2+1].  transpose = 0: components
 * 0, 1 hold (F^_x, Phi^) as mcpm_xfuse_force_peer(transpose = 2) leaves them; on return components 1, 2 hold
 * F^_y = -(i g_y) Phi^, F^_z = -(i g_z) Phi^ (gradient_hat, nbody.py:136-163; Hermitian-consistent).  transpose = 1:
 * (C^_x, C^_y, C^_z) -> component 1 = g_y C^_y + g_z C^_z, the second input of mcpm_xfuse_force_peer(transpose = 3). */
int mcpm_yz_gradients(void* stream, void* buf3_c64, int xl, int ny, int nz, int grad_fd, int transpose);
/* out[0] = max(out[0], max_i |x[i * stride]|), out[0] >= 0 on entry (device float; NaN sticks): the halo guard of a
 * slab-decomposed caller -- the largest displacement of its particles across the slab direction -- as one pass. */
int mcpm_absmax(void* stream, const float* x, int64_t n, int stride, float* out);
int mcpm_rsd_shift(void* stream, const float* pos, const float* vel, const float los[3], float coef, int64_t np,
                   float* pos_out);
int mcpm_rsd_shift_vjp(void* stream, const float* posbar, const float los[3], float coef, int64_t np, float* velbar,
                       int accumulate);

/* ---- Lagrangian bias expansion (montecosmo/bricks.py:327-452) as fused passes -----------------
 * bias_spectra: ONE pass over delta_k [nx,ny,nz/2+1] writing the M = 10 (inv_transfer given: 12) spectra whose inverse
 * transforms the expansion needs, in this order: s00, s11, s01, s02, s12 (s_ij = (k_i k_j / k^2 - delta_ij / 3) delta,
 * bricks.py:361-372), delta, -k^2 delta (:397), i k_x|y|z delta (:444-446), [phi = delta * inv_transfer, -k^2 phi (:411-413,
 * 436)].  cells_per_len[d] = n_d / box_d turns rad/cell into h/Mpc (rfftk with box_size, nbody.py:50-77).  Every
 * output is Hermitian-consistent (feeds a C2R).  _vjp: its transpose, dkbar (+)= sum_m conj(multiplier_m) outbar[m]. */
int mcpm_bias_spectra(void* stream, const void* delta_k_c64, int nx, int ny, int nz, const float cells_per_len[3],
                      const float* inv_transfer, void* out_c64);
int mcpm_bias_spectra_vjp(void* stream, const void* outbar_c64, int nx, int ny, int nz, const float cells_per_len[3],
                          const float* inv_transfer, void* dkbar_c64, int accumulate);
/* s5 = [s00, s11, s01, s02, s12][n] -> out2 = [s^2, s^3][n] with s22 = -(s00 + s11) (bricks.py:373-395), and its VJP. */
int mcpm_shear_invariants(void* stream, const float* s5, int64_t n, float* out2);
int mcpm_shear_invariants_vjp(void* stream, const float* s5, const float* out2bar, int64_t n, float* s5bar);
/* Particle side.  vals [np, K]: the K = 7 (9 with PNG) fields read at each particle, in the order delta, s^2, s^3,
 * lap delta, grad delta (3), [phi, lap phi].  growth: one value (growth_arr = NULL) or one per particle (light cone).
 * coef[13] = b1, b2, bs2, b3, bds2, bs3, bn2, bnpar, fNL_bp, fNL_bpd, fNL_bpd2, fNL_bps2, fNL_bn2p.
 * bias_moments: mom_f64[0] += sum (delta g)^2, mom_f64[1] += sum phi delta g (np times the two means, :354, :421).
 * bias_weights: weights [np] and dvel [np, 3] (nullable) of bricks.py:350-449, reading the moments from device memory.
 * bias_weights_vjp: wbar [np], dvelbar [np, 3] (nullable) -> valsbar [np, K]; coefbar_f64[14] (accumulated): the 13
 * coefficients, then the scalar growth factor; gbar_arr [np] when growth_arr is given.  msum_f64[2]: zeroed scratch. */
int mcpm_bias_moments(void* stream, const float* vals, int K, float growth, const float* growth_arr, int64_t np,
                      double* mom_f64);
int mcpm_bias_weights(void* stream, const float* vals, int K, float growth, const float* growth_arr,
                      const float coef[13], const double* mom_f64, int64_t np, float* weights, float* dvel);
int mcpm_bias_weights_vjp(void* stream, const float* vals, int K, float growth, const float* growth_arr,
                          const float coef[13], const double* mom_f64, const float* wbar, const float* dvelbar,
                          int64_t np, double* msum_f64, float* valsbar, double* coefbar_f64, float* gbar_arr);

/* out = in * t (real transfer, half-spectrum shaped) ; white2lin (bricks.py:152-157) and its transpose */
int mcpm_scale_spectrum(void* stream, const void* in, const float* t, void* out, int64_t nc);

/* ---- particle updates -----------------------------------------------------------------------------------------
 * lpt combine (nbody.py:656-665): dpos = d1*F1 - d2*F2 ; vel = F1 - dv2*F2 ; pos_out = pos + dpos (pos_out may be
 * NULL).  f2 may be NULL (1LPT). */
int mcpm_lpt_combine(void* stream, const float* pos, const float* f1, const float* f2, float d1, float d2,
                     float dv2, int64_t np, float* dpos, float* vel, float* pos_out);

/* BullFrog kick fused with the force readout and the following drift (nbody.py:933-951):
 *   F = read(pos, fmesh[3]) ; vel = alpha*vel + beta*F ; pos += vel * drift.
 * In place on pos / vel.  force_out may be NULL. */
int mcpm_kick_drift(void* stream, float* pos, float* vel, const float* fmesh3, int64_t np, int nx, int ny, int nz,
                    int order, float alpha, float beta, float drift, float* force_out);
/* ---- CIC step kernels on the float4-interleaved vector mesh "mesh4" = [nx, ny, nz] cells of {c0, c1, c2, c3} -------
 * These are what mcpm_nbody_steps / _vjp run for order 2; exposed so that they can be timed and tested in isolation.
 * interleave3 / deinterleave3: three planar meshes <-> mesh4.xyz (w = 0), the conversion around cuFFT.
 * kick_drift4: as mcpm_kick_drift with the forces read from mesh4.xyz (8 x 16-byte loads per particle).
 * paint3v4 (reverse step, first half): vbar += xbar * drift (stored, when xbar != NULL);
 *           mesh4[cell].xyz += scale * vbar * W over the 8 CIC corners (8 x red.global.add.v4.f32). mesh4 is NOT zeroed.
 * read_grad4v (reverse step, second half): xbar += d/dx [ cscale * vbar . F(x) + rhobar(x) ], F = mesh4.xyz, rhobar
 *           planar; then vbar *= alpha. */
int mcpm_interleave3(void* stream, const float* planar3, float* mesh4, int64_t n);
int mcpm_deinterleave3(void* stream, const float* mesh4, float* planar3, int64_t n);
int mcpm_kick_drift4(void* stream, float* pos, float* vel, const float* fmesh4, int64_t np, int nx, int ny, int nz,
                     float alpha, float beta, float drift);
int mcpm_paint3v4(void* stream, const float* pos, float* vbar, const float* xbar, float drift, float scale, int64_t np,
                  int nx, int ny, int nz, float* mesh4);
int mcpm_read_grad4v(void* stream, const float* pos, const float* fmesh4, const float* rhobar, float* vbar,
                     float cscale, float alpha, int64_t np, int nx, int ny, int nz, float* xbar);

/* ---- position frames for the stateless particle kernels -------------------------------------------------------
 * relative = 0: absolute positions (as every entry point above).  relative = 1: pos[p] is the displacement from
 * lattice site  o_a + i_a * s_a / p_a  (target-mesh cells), (i_x, i_y, i_z) = unravel(p, (px, py, pz)); px*py*pz must
 * equal np.  (sx, sy, sz) = the mesh cells the lattice spans (regular_pos: the mesh shape; a slab-decomposed rank:
 * its own xl x ny x nz with ox = halo planes).  scale[] multiplies the displacement only; the site is already in
 * target-mesh units.  A NULL frame means absolute.  The brick variants need a spacing of exactly one cell. */
typedef struct mcpm_frame {
  int relative;
  int px, py, pz;
  int ox, oy, oz;
  int sx, sy, sz;
} mcpm_frame;
int mcpm_paint_f(void* stream, const mcpm_frame* frame, const float* pos, const float* weights, float wscalar,
                 int64_t np, int nx, int ny, int nz, int order, const float scale[3], float shift, float* mesh,
                 int accumulate);
int mcpm_read_f(void* stream, const mcpm_frame* frame, const float* pos, const float* mesh, int nmesh, int64_t np,
                int nx, int ny, int nz, int order, const float scale[3], float shift, float* out);
int mcpm_read_grad_f(void* stream, const mcpm_frame* frame, const float* pos, const float* mesh, int nmesh,
                     const float* cot, int64_t np, int nx, int ny, int nz, int order, const float scale[3],
                     float shift, float* grad, int accumulate);
int mcpm_paint_vjp_f(void* stream, const mcpm_frame* frame, const float* pos, const float* weights, float wscalar,
                     const float* mesh_bar, int64_t np, int nx, int ny, int nz, int order, const float scale[3],
                     float shift, float* posbar, float* weightsbar, int accumulate);
int mcpm_paint3_f(void* stream, const mcpm_frame* frame, const float* pos, const float* vals3, float vscale, int64_t np,
                  int nx, int ny, int nz, int order, float* mesh3, int accumulate);
int mcpm_kick_drift_f(void* stream, const mcpm_frame* frame, float* pos, float* vel, const float* fmesh3, int64_t np,
                      int nx, int ny, int nz, int order, float alpha, float beta, float drift, float* force_out);
int mcpm_kick_drift4_f(void* stream, const mcpm_frame* frame, float* pos, float* vel, const float* fmesh4, int64_t np,
                       int nx, int ny, int nz, float alpha, float beta, float drift);
int mcpm_paint3v4_f(void* stream, const mcpm_frame* frame, const float* pos, float* vbar, const float* xbar,
                    float drift, float scale, int64_t np, int nx, int ny, int nz, float* mesh4);
int mcpm_read_grad4v_f(void* stream, const mcpm_frame* frame, const float* pos, const float* fmesh4,
                       const float* rhobar, float* vbar, float cscale, float alpha, int64_t np, int nx, int ny, int nz,
                       float* xbar);
/* read_grad4v with the tail of the reverse step fused in: xbar += d/dx [...], then vbar = alpha * vbar + dnext * xbar
 * (the kick's vbar *= alpha and the NEXT reverse step's leading vbar += xbar * drift while both are in registers; the
 * scatter of that next step then reads vbar only: mcpm_paint3_brick with xbar = NULL).  What mcpm_nbody_steps_vjp runs. */
int mcpm_read_grad4v_step_f(void* stream, const mcpm_frame* frame, const float* pos, const float* fmesh4,
                            const float* rhobar, float* vbar, float cscale, float alpha, float dnext, int64_t np, int nx,
                            int ny, int nz, float* xbar);
int mcpm_paint_brick_f(void* stream, const mcpm_frame* frame, int px, int py, int pz, const float* pos,
                       const float* weights, float wscalar, float shift, int64_t np, int nx, int ny, int nz,
                       float* mesh);
int mcpm_paint3_brick_f(void* stream, const mcpm_frame* frame, int px, int py, int pz, const float* pos, float* vbar,
                        const float* xbar, float drift, float scale, int64_t np, int nx, int ny, int nz,
                        float* mesh3);

/* pos += vel * drift */
int mcpm_drift(void* stream, float* pos, const float* vel, float drift, int64_t np);

/* ---- composite operators (same names as montecosmo/nbody.py) --------------------------------------------------
 * pm_forces with a painted mesh (nbody.py:583-604, mesh given as a shape tuple):
 * fmesh3 receives the three real force meshes [3, nx, ny, nz]; forces [np, 3] may be NULL. */
int mcpm_pm_forces(mcpm_engine* eng, void* stream, const float* pos, int64_t np, int order, int paint_deconv,
                   int lap_fd, int grad_fd, float kcut, float* fmesh3, float* forces);
/* its VJP w.r.t. pos: fbar [np,3], fmesh3 from the forward (or recomputed by it) -> posbar [np,3] (accumulated
 * when accumulate != 0). */
int mcpm_pm_forces_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* fbar, const float* fmesh3,
                       int64_t np, int order, int paint_deconv, int lap_fd, int grad_fd, float kcut,
                       float* posbar, int accumulate);

/* pm_forces with a given spectrum (nbody.py:595-604) and pm_forces2 (nbody.py:607-631); read order `order`. */
int mcpm_pm_forces_mesh(mcpm_engine* eng, void* stream, const float* pos, const void* delta_k, int64_t np,
                        int order, int lap_fd, int grad_fd, float kcut, float* forces);
int mcpm_pm_forces2(mcpm_engine* eng, void* stream, const float* pos, const void* delta_k, int64_t np, int order,
                    int lap_fd, int grad_fd, float* forces, float* h6_out /* nullable: [6,nx,ny,nz] tape */);

/* lpt (nbody.py:634-667), scalar scale factor.  Growth coefficients are host scalars computed by the caller
 * (d1 = a2g(a), d2 = a2g2(a), dv2 = a2dg2dg(a), nbody.py:750-777) so they stay differentiable on the host.
 * Outputs dpos, vel [np,3]; tape: f1 [np,3], f2 [np,3], h6 [6,nx,ny,nz] (each nullable when no VJP is wanted).
 * pos == NULL (mcpm_lpt, mcpm_lpt_vjp, mcpm_pm_forces_mesh, mcpm_pm_forces2): the particles sit on the cells of the mesh,
 * pos = regular_pos(mesh_shape) (bricks.py:593-603; np = nx*ny*nz), which is how nbody_bf and model.evolve call lpt
 * (nbody.py:984, model.py:763).  NGP and CIC reads there return the cell value itself, so the three reads and their
 * transposes become layout changes instead of gathers / scatters. */
int mcpm_lpt(mcpm_engine* eng, void* stream, const void* delta_k, const float* pos, int64_t np, int lpt_order,
             int read_order, int lap_fd, int grad_fd, float d1, float d2, float dv2, float* dpos, float* vel,
             float* f1, float* f2, float* h6);
/* VJP of lpt w.r.t. delta_k and the three coefficients: dposbar, velbar -> dkbar (complex, convention above),
 * coefbar[3] (device, float64: d1bar, d2bar, dv2bar). */
int mcpm_lpt_vjp(mcpm_engine* eng, void* stream, const float* pos, int64_t np, int lpt_order, int read_order,
                 int lap_fd, int grad_fd, float d1, float d2, float dv2, const float* dposbar, const float* velbar,
                 const float* f1, const float* f2, const float* h6, void* dkbar, double* coefbar, int accumulate);

/* nbody_bf step loop (nbody.py:933-951, 999): n_steps drift-kick-drift steps in place on (pos, vel).
 * Per-step host arrays: alpha[s], beta[s] = (1 - alpha) / g1, drift_pre[s], drift_post[s] (= dg/2 each).
 * Tape (nullable): xk [n_steps, np, 3] positions at kick time, vk [n_steps, np, 3] velocities after the kick,
 * fm [n_steps, 4, nx, ny, nz] floats of force meshes in an engine-defined layout (float4-interleaved for CIC). */
int mcpm_nbody_steps(mcpm_engine* eng, void* stream, float* pos, float* vel, int64_t np, int n_steps,
                     const float* alpha, const float* beta, const float* drift_pre, const float* drift_post,
                     int order, int paint_deconv, int lap_fd, int grad_fd, float* xk, float* vk, float* fm);
/* VJP of the loop: (posbar, velbar) at the end -> at the start, in place.  v0 = velocities before the first step.
 * coefbar (device float64 [n_steps, 4]: alpha, beta, drift_pre, drift_post cotangents) may be NULL.
 * fm may be NULL (pass fm = NULL to mcpm_nbody_steps too): the force meshes are then not taped -- 16 bytes per cell and
 * step -- and every reverse step recomputes its own from xk[s] (one extra paint + forward Fourier pass per step), the
 * memory / recompute trade of the reference's checkpointed adjoint (diffrax, nbody.py:999). */
int mcpm_nbody_steps_vjp(mcpm_engine* eng, void* stream, float* posbar, float* velbar, int64_t np, int n_steps,
                         const float* alpha, const float* beta, const float* drift_pre, const float* drift_post,
                         int order, int paint_deconv, int lap_fd, int grad_fd, const float* xk, const float* vk,
                         const float* fm, const float* v0, double* coefbar);

/* nufft (nbody.py:532-577) for paint_shape == engine shape: interlaced, Jacobian-scaled, deconvolved paint returning
 * the half spectrum at the paint shape (the caller applies mcpm_chreshape to reach final_shape).
 * pos is in FINAL units; scale[d] = paint_shape[d] / final_shape[d]. */
int mcpm_nufft(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar, int64_t np,
               const float scale[3], int paint_order, int interlace_order, int paint_deconv, void* out_k);
/* VJP: outbar_k -> posbar [np,3] and weightsbar [np] (each nullable). */
int mcpm_nufft_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                   int64_t np, const float scale[3], int paint_order, int interlace_order, int paint_deconv,
                   const void* outbar_k, float* posbar, float* weightsbar);

/* nufft of REDSHIFT-SPACE positions with the flat-sky shift applied inside the paint kernels (model.py:780-809 with
 * bricks.py:781-792 in cell units): every particle is deposited at pos + (vel . los) * coef * los, which is never
 * written to memory.  _vjp additionally returns velbar [np,3] = coef * (xbar . los) * los (posbar is the cotangent of
 * pos itself, as in mcpm_nufft_vjp). */
int mcpm_nufft_rsd(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const float los[3], float coef,
                   const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order,
                   int interlace_order, int paint_deconv, void* out_k);
int mcpm_nufft_rsd_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const float los[3], float coef,
                       const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order,
                       int interlace_order, int paint_deconv, const void* outbar_k, float* posbar, float* velbar,
                       float* weightsbar);

/* nufft of OBSERVED positions with the general observation chain of model.py:780-799 applied inside the paint kernels:
 * cell -> physical (bricks.py:628-636) -> line of sight and scale factor (:747-766) -> redshift-space distortion
 * (:781-792, with the velocity bias `dvel`) -> Alcock-Paczynski (ap_auto :795-813 | ap_param :847-856) -> cell
 * (:638-646), for a box placed and rotated with respect to the observer, curved or flat sky, a scalar a_obs or the
 * light cone.  The transformed positions are never written.  Everything that depends on the cosmology comes in as
 * numbers the caller computes from it: `gf` = D(a_obs) f(a_obs), or -- lightcone -- tab_gf[k] = (D f)(a(r_k)), and
 * -- ap == 1 -- tab_ap[k] = chi_fid(a(r_k)) / r_k - 1, on the uniform radius grid r_k = r0 + k * dr, k < nt (device
 * arrays, linearly interpolated, clamped at the ends).
 *   cell[d]   Mpc/h per unit of pos;   origin = R^T box_center - box_size / 2;   los = R^T box_center / |box_center|
 *   rot       R (row-major), used for `dvel` [np,3] (observer frame, Mpc/h, nullable) only
 *   ap        0 none | 1 table | 2 parametrised: a_par, a_perp (curved sky: a_par = alpha_iso)
 *   rsd       0: no redshift-space term (vel may be NULL)
 * _vjp returns posbar, velbar, dvelbar, weightsbar (each nullable) and parbar (nullable), float64
 * [MCPM_OBS_SLOTS][3 + 2 nt] of which ROW 0 holds the result on return: cotangents of gf, a_par, a_perp, tab_gf[nt],
 * tab_ap[nt] (the other rows are scratch for the accumulation).  kcut > 0: Kaiser-Bessel window of that cut. */
#define MCPM_OBS_SLOTS 32
typedef struct mcpm_obs {
  int curved, lightcone, ap, rsd;
  float cell[3], origin[3], los[3];
  float gf, a_par, a_perp;
  float r0, dr;
  int nt;
  const float* tab_gf;
  const float* tab_ap;
  const float* dvel;
  float rot[9];
  const float* par; /* device [3] = gf, a_par, a_perp replacing the three host values above (a caller whose scalars are
                       traced device values, e.g. under XLA); NULL: the host values */
} mcpm_obs;
int mcpm_nufft_obs(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const mcpm_obs* obs,
                   const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order, float kcut,
                   int interlace_order, int paint_deconv, void* out_k);
int mcpm_nufft_obs_vjp(mcpm_engine* eng, void* stream, const float* pos, const float* vel, const mcpm_obs* obs,
                       const float* weights, float wscalar, int64_t np, const float scale[3], int paint_order, float kcut,
                       int interlace_order, int paint_deconv, const void* outbar_k, float* posbar, float* velbar,
                       float* dvelbar, float* weightsbar, double* parbar);

/* Functions of the comoving distance evaluated at the particles -- the light cone: los_scalefactor_pos
 * (bricks.py:747-766: r = |cell2phys_pos(pos)| on a curved sky, |l . cell2phys_pos(pos)| on a flat one) followed by the
 * growth lookups a2g, a2g2, a2f, ... of lagrangian_bias (bricks.py:341) and lpt (nbody.py:651-665) -- as ONE pass:
 * out[p, k] = tab_k(r_p), tabs = device [ntab][nt] on the radius grid of `geom` (r0, dr, nt), which also supplies
 * curved, cell, origin and los (the other fields are ignored).  _vjp: posbar [np,3] (nullable) and tabbar (nullable),
 * float64 [MCPM_OBS_SLOTS][ntab][nt] whose first [ntab][nt] block holds the cotangent of the tables on return. */
int mcpm_radial_tables(void* stream, const float* pos, int64_t np, const mcpm_obs* geom, int ntab, const float* tabs,
                       float* out);
int mcpm_radial_tables_vjp(void* stream, const float* pos, int64_t np, const mcpm_obs* geom, int ntab, const float* tabs,
                           const float* outbar, float* posbar, double* tabbar);

/* nufft / its VJP with kernel_type = 'kaiser_bessel' (nbody.py:532-577 with the window of 280-312). */
int mcpm_nufft_kb(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar, int64_t np,
                  const float scale[3], int paint_order, float kcut, int interlace_order, int paint_deconv,
                  void* out_k);
int mcpm_nufft_vjp_kb(mcpm_engine* eng, void* stream, const float* pos, const float* weights, float wscalar,
                      int64_t np, const float scale[3], int paint_order, float kcut, int interlace_order,
                      int paint_deconv, const void* outbar_k, float* posbar, float* weightsbar);

#ifdef __cplusplus
}
#endif
#endif /* MCPM_H_ */
