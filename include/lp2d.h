/* lp2d.h — C ABI of liblp2d.so, the B200 (sm_100a) backend of the lightprop2d
 * angular-spectrum propagation hot path.
 *
 * The reference (vongostev/lightprop2d, pure Python) has no FFI; its backend seam is
 * the array module `self.xp` plus `_fft2/_ifft2` and `_np/_asxp` inside `Beam2D`
 * (lightprop2d/beam2d.py:83, 105-114, 135-187, 207-229).  Each entry point below
 * replaces the numpy/cupy calls of one Beam2D method; the line numbers cite the
 * reference code that the call reproduces.  The binding a maintainer would add to
 * the reference is a ctypes stub, shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every array argument is a DEVICE pointer owned by the
 *     caller (row-major, C-contiguous, beams stacked along the leading axis) unless
 *     the name starts with `host_`;
 *   - complex64 = interleaved float pairs, complex128 = interleaved double pairs;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); every call
 *     only enqueues work, nothing synchronises;
 *   - return value 0 = success, otherwise an LP2D_E* code; lp2d_last_error() gives
 *     the message of the last failure on the calling thread;
 *   - the plan is bound to the device that was current when it was created.
 */
#ifndef LP2D_H
#define LP2D_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LP2D_API __attribute__((visibility("default")))
#else
#define LP2D_API
#endif

#define LP2D_VERSION 100

#define LP2D_OK 0
#define LP2D_EINVAL 1       /* bad argument */
#define LP2D_EUNSUPPORTED 2 /* grid size / combination not built */
#define LP2D_ECUDA 3        /* CUDA runtime error (message in lp2d_last_error) */

#define LP2D_C64 0
#define LP2D_C128 1

/* mask element types accepted by lp2d_mul_fft2 */
#define LP2D_MASK_COMPLEX 1 /* same complex type as the field */
#define LP2D_MASK_REAL 2    /* real type of the field (float / double) */
#define LP2D_MASK_U8 3      /* bytes, non-zero = 1 (numpy bool) */

typedef struct lp2d_plan lp2d_plan;

LP2D_API int lp2d_version(void);
LP2D_API const char* lp2d_last_error(void);

/* Smallest / largest supported npoints (powers of two in between are supported). */
LP2D_API int lp2d_min_npoints(void);
LP2D_API int lp2d_max_npoints(int dtype);
/* Longest line lp2d_pass_lines transforms.  Plans of lp2d_max_npoints < n <= this hold line kernels
 * only (one line spread over a thread-block cluster): every 2-D entry point returns
 * LP2D_EUNSUPPORTED for them and the grid is driven as a slab decomposition (slab.py), also on
 * one GPU. */
LP2D_API int lp2d_max_line_points(int dtype);

/* Host helper: the grids of Beam2D._construct_grids (beam2d.py:299-308) with the
 * same IEEE operation order: dL = area/n; X[k] = (k - ceil(n/2)) * dL for the
 * arange(-n//2, n//2) of the reference; Kx = fftfreq(n, dL) = idx * (1/(n*dL));
 * a[i] = k0*k0 - Kx[i]^2, b[j] = Kx[j]^2 (so Kz[j][i] = sqrt(a[i] - b[j]), NaN -> 0);
 * x2[i] = X[i]^2.  Any output pointer may be NULL.  All arrays are HOST, length n. */
LP2D_API int lp2d_grid_tables(int n, double area_size, double wl, double* host_X, double* host_Kx,
                     double* host_a, double* host_b, double* host_x2);

/* Plan: twiddle tables and the three FP64 tables (copied from host) for an n x n
 * grid of the given complex type.  Replaces the per-object grid state of
 * Beam2D.__post_init__ / _construct_grids (beam2d.py:93-133, 290-308). */
LP2D_API int lp2d_plan_create(lp2d_plan** plan, int n, int dtype, const double* host_a,
                     const double* host_b, const double* host_x2);
LP2D_API int lp2d_plan_destroy(lp2d_plan* plan);

/* _fft2 / _ifft2 (beam2d.py:207-229): 2-D DFT over the last two axes of `batch`
 * planes; forward unscaled, inverse scaled by 1/n^2.  `out` may equal `in`. */
LP2D_API int lp2d_fft2(lp2d_plan* plan, const void* in, void* out, long long batch, int inverse,
              void* stream);

/* Beam2D.propagate, stateful form (beam2d.py:481-490): spectrum *= H(z) in place,
 * field = ifft2(spectrum).  `intensity` (real, may be NULL) receives |field|^2
 * (iprofile, beam2d.py:717) from the same pass.  z == 0 is the caller's early
 * return. */
LP2D_API int lp2d_propagate(lp2d_plan* plan, void* spectrum, void* field, void* intensity,
                   long long batch, double z, void* stream);

/* A z-sweep in ONE call: nsteps x [propagate(dz)] as lp2d_propagate, and after every step
 *   probe_out[k][b]  (device, real type, may be NULL) = |field_b[probe_index]|^2, the number the reference's sweep
 *                    loops read each step (tests/test_analytics.py:43-50, tests/hole_diffraction.py:37-42);
 *   moments_out[k][b][5] (device, FP64, may be NULL) = [sum I, sum x I, sum y I, sum x^2 I, sum y^2 I] about the
 *                    grid origin (x = (col - n/2) dL), from which centroid and D4sigma of every step follow
 *                    (beam2d.py:680-684, 696-702) without a device-to-host round trip per step.
 * Everything is enqueued on `stream`; the caller synchronises once, after the sweep. */
LP2D_API int lp2d_propagate_sweep(lp2d_plan* plan, void* spectrum, void* field, long long batch, int nsteps,
                         double dz, long long probe_index, void* probe_out, double dL, double* moments_out,
                         void* stream);

/* Fresh chain for ensembles: field_out = ifft2(fft2(field_in) * H(z)) in three
 * passes without materialising the spectrum (unless spectrum_out != NULL, which
 * receives fft2(field_in) * H(z), the spectrum state the reference would hold).
 * field_out may be NULL when only `intensity` is wanted; then `work` (>= one plane
 * of complex samples, work_bytes long) carries the intermediates.  field_out may
 * equal field_in. */
LP2D_API int lp2d_propagate_field(lp2d_plan* plan, const void* field_in, void* field_out,
                         void* spectrum_out, void* intensity, void* work,
                         long long work_bytes, long long batch, double z, void* stream);

/* Beam2D.lens (beam2d.py:508-511): field *= exp(i (X^2+Y^2) k0 / (2f)) in place
 * (k0 = 1/wl, as the reference), spectrum = fft2(field). */
LP2D_API int lp2d_lens(lp2d_plan* plan, void* field, void* spectrum, long long batch, double k0,
              double f, void* stream);

/* coordinate_filter / spectral_filter with an array mask (beam2d.py:338-342,
 * 369-373): data *= mask in place (mask NULL = no multiply), then
 * other = fft2(data) (inverse = 0) or ifft2(data) (inverse = 1).
 * mask_beam_stride = 0 broadcasts one n x n mask over the batch. */
LP2D_API int lp2d_mul_fft2(lp2d_plan* plan, void* data, void* other, const void* mask, int mask_kind,
                  long long mask_beam_stride, long long batch, int inverse, void* stream);

/* One pass of the decomposition, for callers that interleave their own step between
 * the row and the column pass (the slab-decomposed grid exchanges tiles in between)
 * and for per-kernel timing.  axis: rows are the contiguous axis.
 *   LP2D_OP_FFT        out = (i)fft along the axis (inverse scaled by 1/n)
 *   LP2D_OP_H_FFT      rows only: data *= H(z); out_state (may be NULL) = data*H;
 *                      out = (i)fft of that along the rows
 *   LP2D_OP_FFT_H_IFFT columns only: fft, *= H(z), out_state (may be NULL) = that,
 *                      out = ifft (scaled by 1/n)
 * `intensity` (may be NULL) = |out|^2.  out may be NULL when only intensity is wanted. */
#define LP2D_AXIS_ROWS 0
#define LP2D_AXIS_COLS 1
#define LP2D_OP_FFT 0
#define LP2D_OP_H_FFT 1
#define LP2D_OP_FFT_H_IFFT 2
LP2D_API int lp2d_pass(lp2d_plan* plan, int axis, int op, const void* in, void* out, void* out_state,
              void* intensity, long long batch, int inverse, double z, void* stream);

/* Row pass over a SLAB of one plane: lines line0 .. line0 + nlines - 1 of an n x n plane, for the
 * slab-decomposed grid (one row slab per GPU, all-to-all transpose between the two passes; the
 * second pass runs on the transposed slab with a plan built from lp2d_grid_tables' a/b tables
 * swapped and negated, see lightprop2d_b200/slab.py).  `in`, `out`, `out_state`, `intensity` point
 * at the first line OF THE SLAB (line0), not of the plane.  op: LP2D_OP_FFT or LP2D_OP_H_FFT. */
LP2D_API int lp2d_pass_lines(lp2d_plan* plan, int op, const void* in, void* out, void* out_state,
                    void* intensity, int line0, int nlines, int inverse, double z, void* stream);

/* The same pass reading its lines straight from the receive buffer of the tile exchange, so that no
 * unpack copy is needed: `in_tiles` holds n / tile tiles, tile c = (nlines x tile) samples row-major,
 * and sample pos of slab line i is in_tiles[(pos / tile) * nlines * tile + i * tile + pos % tile] (what
 * all_to_all of lp2d_pack_transpose output delivers, tile = n / ranks).  tile: a power of two dividing
 * n.  out / out_state / intensity are contiguous slab lines as in lp2d_pass_lines and must not alias
 * in_tiles. */
LP2D_API int lp2d_pass_lines_tiles(lp2d_plan* plan, int op, const void* in_tiles, int tile, void* out,
                          void* out_state, void* intensity, int line0, int nlines, int inverse,
                          double z, void* stream);

/* Transposing tile pack for the slab exchange: in is a (rows x ld) row-major slab; for every tile
 * c (0 .. ntiles-1) of `tile` columns, out[c] (tile x rows, contiguous) = transpose(in[:, c*tile :
 * (c+1)*tile]).  ntiles * tile == ld. */
LP2D_API int lp2d_pack_transpose(int dtype, const void* in, void* out, int rows, int ld, int tile,
                        void* stream);

/* iprofile (beam2d.py:717): out = |field|^2 (float for complex64, double for complex128). */
LP2D_API int lp2d_intensity(lp2d_plan* plan, const void* field, void* out, long long batch, void* stream);

/* Beam2D.expand (beam2d.py:392-401): edge-replicating pad of every n_in x n_in plane to
 * n_out x n_out ((n_out - n_in) even, the same pad on every side).  No plan needed. */
LP2D_API int lp2d_expand(int dtype, const void* in, void* out, int n_in, int n_out, long long batch,
                void* stream);

/* The generators of lightprop2d/profiles.py:35-136 evaluated on the device: `batch`
 * planes of n x n samples from the coordinate tables dev_x / dev_y (device, n doubles each: Beam2D.X and Beam2D.Y --
 * where the reference shifts them in place, profiles.py:69-70, the caller does so before the upload).
 * out_kind: LP2D_C64 / LP2D_C128 (complex field) or LP2D_GEN_OUT_MASK (one byte per sample; 0/1 profiles only).
 * host_params6 per kind:
 *   PLANE_WAVE -;  RANDOM_WAVE, RANDOM_WAVE_BIN -;  GAUSSIAN_BEAM {A0, rho0};
 *   ROUND_HOLE, RANDOM_ROUND_HOLE(_BIN, _PHASE) {R, x0, y0, -, -, 1/exp(0.5) as the host evaluates it};
 *   RECTANGLE_HOLE {dx, dy, x0, y0} (square_hole: dx = dy);  SQUARE_SLITS {d, slits_distance, x0, y0}.
 * Deterministic profiles use the reference's FP64 expressions in its operation order (device exp() may differ from
 * the host's in the last bit: a sample whose Gaussian value is within one ulp of the round_hole threshold can fall
 * on the other side).  Random profiles draw one uniform number per cell of binning_order x binning_order samples
 * from Philox4x32-10 keyed by (seed, first_beam + b, cell): the numbers lp2d_random_fields draws, independent of
 * batch and rank; numpy's stream is not reproduced (seeded parity uses the host generators). */
#define LP2D_GEN_PLANE_WAVE 0
#define LP2D_GEN_RANDOM_WAVE 1
#define LP2D_GEN_RANDOM_WAVE_BIN 2
#define LP2D_GEN_GAUSSIAN_BEAM 3
#define LP2D_GEN_ROUND_HOLE 4
#define LP2D_GEN_RANDOM_ROUND_HOLE 5
#define LP2D_GEN_RANDOM_ROUND_HOLE_BIN 6
#define LP2D_GEN_RANDOM_ROUND_HOLE_PHASE 7
#define LP2D_GEN_RECTANGLE_HOLE 8
#define LP2D_GEN_SQUARE_SLITS 9
#define LP2D_GEN_OUT_MASK 16
LP2D_API int lp2d_generate(int out_kind, void* out, int n, long long batch, const double* dev_x, const double* dev_y,
                           int kind, const double* host_params6, int binning_order, unsigned long long seed,
                           unsigned long long first_beam, void* stream);

/* Beam2D.coarse (beam2d.py:456-459): mean over m x m blocks, n_in x n_in -> n_in/m x n_in/m. */
LP2D_API int lp2d_coarse(int dtype, const void* in, void* out, int n_in, int m, long long batch,
                void* stream);

/* Ensemble inputs generated on the device: `batch` fields of
 * profiles.random_round_hole (kind 0: amplitude uniform in [0,1) inside the aperture, profiles.py:94-101)
 * or profiles.random_round_hole_phase (kind 1: exp(2 pi i u) inside, profiles.py:104-113), zero outside.
 * `mask_u8` (device, n*n bytes) is the host-built profiles.round_hole array, so the aperture is exactly
 * the reference's.  The uniform numbers come from a counter-based generator (Philox4x32-10) keyed by
 * (seed, first_beam + b, sample index): a beam does not depend on the batch or the rank it is generated in
 * (the numpy stream of the reference cannot be reproduced on the device; parity tests use host inputs). */
LP2D_API int lp2d_random_fields(int dtype, void* out, const void* mask_u8, int n, long long batch,
                       unsigned long long first_beam, unsigned long long seed, int kind, void* stream);

/* phiprofile (beam2d.py:734): out = atan2(im, re) for `count` samples. */
LP2D_API int lp2d_phase(int dtype, const void* field, void* out, long long count, void* stream);

/* centroid / D4sigma sums (beam2d.py:680-684, 696-702) of ONE n x n field, FP64:
 * out5 (device) = [sum I, sum x I, sum y I, sum (x-xc)^2 I, sum (y-yc)^2 I],
 * x = (col - n/2) dL, y = (row - n/2) dL, I = |field|^2 in the field's real type. */
LP2D_API int lp2d_moments(int dtype, const void* field, int n, double dL, double xc, double yc,
                 double* out5, void* stream);

/* deconstruct_by_modes(mode='scalar') (beam2d.py:592-600): coeffs[b][k] =
 * sum_p modes[k][p] * field_b[centre nb x nb crop][p]  (no conjugation).
 * modes: nmodes x nb*nb, coeffs: batch x nmodes, same complex type as the field. */
LP2D_API int lp2d_modes_project(int dtype, const void* field, const void* modes, void* coeffs, int n,
                       int nb, int nmodes, long long batch, void* stream);

/* The same projection for a strip of the crop held by one rank of a slab-decomposed grid (SURVEY.md section 8e
 * row 3): coeffs[k] = sum_{p < count} modes[k * ldm + p] * flat_field[p].  The caller adds the partial vectors of
 * the ranks (one all-reduce of nmodes complex numbers, lightprop2d_b200/slab.py). */
LP2D_API int lp2d_modes_project_flat(int dtype, const void* flat_field, const void* modes, void* coeffs,
                            long long count, long long ldm, int nmodes, void* stream);

/* construct_by_modes (beam2d.py:552-562, 664-666): field_b = sum_k coeffs[b][k] * mode_k
 * zero-padded from nb x nb to the centred window of the n x n grid. */
LP2D_API int lp2d_modes_synthesize(int dtype, const void* coeffs, const void* modes, void* field, int n,
                          int nb, int nmodes, long long batch, void* stream);

/* ---- grids the power-of-two kernel sets do not cover (any npoints; grids above lp2d_max_npoints) ----
 * The reference accepts every npoints (beam2d.py:299-308) and expand() produces e.g. 82 points
 * (beam2d.py:385-389).  Such grids run as: pointwise factor kernels + a 2-D DFT by Bluestein's algorithm on
 * top of the power-of-two lp2d_fft2 (lightprop2d_b200/generic.py), all on the device. */

/* data[b][r][c] *= factor(row0 + r, c) on `batch` slabs of `rows` rows of an n x n plane (any n), in place:
 *   LP2D_PW_H    exp(i 2 pi (trunc(Kz z) - Kz z)), Kz^2 = a[c] - b[row0 + r], p0 = z      (beam2d.py:485-489)
 *   LP2D_PW_LENS exp(i (x2[c] + x2[row0 + r]) k0 / (2 f)), p0 = k0, p1 = 2 f              (beam2d.py:508-510)
 * dev_tabs: DEVICE, 3 n doubles a | b | x2 (lp2d_grid_tables).  FP64 phase, the reference's operation order. */
#define LP2D_PW_H 1
#define LP2D_PW_LENS 2
LP2D_API int lp2d_pointwise(int dtype, void* data, long long batch, int rows, int n, int row0, int kind,
                   const double* dev_tabs, double p0, double p1, void* stream);

/* data[b][i] *= mask[b * mask_beam_stride + i], i < plane (beam2d.py:338, 369); mask_kind as lp2d_mul_fft2. */
LP2D_API int lp2d_mul_mask(int dtype, void* data, const void* mask, int mask_kind, long long plane,
                  long long mask_beam_stride, long long batch, void* stream);

/* iprofile without a plan: out = |field|^2 for `count` samples (beam2d.py:717). */
LP2D_API int lp2d_intensity_n(int dtype, const void* field, void* out, long long count, void* stream);

/* Bluestein: X = post(ifft2_m(fft2_m(pre(x)) * fft2_m(filter))) is the n x n DFT (forward unscaled, inverse
 * scaled by 1/n^2, _fft2 / _ifft2 of beam2d.py:207-229) for ANY n, with m = lp2d_bluestein_size(n) a power of
 * two >= 2n - 1.  pre: (batch, n, n) -> (batch, m, m) chirp multiply + zero pad; filter: the (m, m) wrapped
 * conjugate chirp (transform it once with lp2d_fft2); post: chirp multiply + crop (+ 1/n^2). */
LP2D_API int lp2d_bluestein_size(int n);
LP2D_API int lp2d_bluestein_pre(int dtype, const void* in, void* out, int n, int m, long long batch, int inverse,
                       void* stream);
LP2D_API int lp2d_bluestein_post(int dtype, const void* in, void* out, int n, int m, long long batch, int inverse,
                        void* stream);
LP2D_API int lp2d_bluestein_filter(int dtype, void* out, int n, int m, int inverse, void* stream);

/* Dense solve A X = B on the device for the mode decomposition (beam2d.py:601-603 through the normal equations,
 * :649-650): one m x m system, nrhs right-hand sides, complex128 row-major, LU with partial pivoting in one CTA.
 * A is destroyed, B receives X.  dev_info (device int): 0, or k + 1 when pivot k is exactly zero (singular). */
LP2D_API int lp2d_solve(void* A, void* B, int m, int nrhs, int* dev_info, void* stream);

/* Tensor-product linear interpolation, the evaluation behind the reference's interp2d (crop, lens_image:
 * beam2d.py:427-434, 527-535): out[b][r][c] = sum_{i,j in {0,1}} in[b][iy[r] + j][ix[c] + i] * wx[i][c] * wy[j][r].
 * `in` planes are in_plane samples apart with row stride ld_in (a window of a larger plane is fine); ix / iy (n_out
 * ints) are the lower knot of the interval, wx / wy (2 x n_out doubles) the two weights -- computed on the host from
 * the coordinate vectors with the reference's expressions (2 n_out numbers), see beam2d.py here. */
LP2D_API int lp2d_resample(int dtype, const void* in, long long in_plane, int ld_in, void* out, int n_out,
                  const int* dev_ix, const double* dev_wx, const int* dev_iy, const double* dev_wy,
                  long long batch, void* stream);

/* Which kernel family serves the steps of this plan that read and write state (propagate, lens,
 * mask + fft2): 0 = one kernel per pass (row pass, column pass), 1 = the fused step kernel (one
 * launch per step, row <-> column intermediate in an L2-resident ring).  `name` (may be NULL)
 * receives a short description (at most name_len - 1 characters). */
LP2D_API int lp2d_plan_step_kernel(lp2d_plan* plan, char* name, int name_len);

/* Spectrum layout of a plan.  transposed = 1: every SPECTRUM buffer the 2-D calls of this plan read or write
 * (lp2d_fft2 output / inverse input, lp2d_propagate / _sweep state, lp2d_lens output, lp2d_mul_fft2 forward output /
 * inverse data, lp2d_propagate_field spectrum_out) is stored transposed: memory row j holds grid column j of the
 * spectrum the reference keeps (Beam2D.spectrum[i, j], beam2d.py:130-133).  Fields, masks of fields and intensities
 * keep their natural orientation; a mask applied to a spectrum (lp2d_mul_fft2 with inverse = 1) must be given
 * transposed too.  Why: the final column stage of a 2-D transform then stores whole rows of the destination
 * instead of 64-byte pieces one row apart, and every HBM access of a step is a contiguous row.
 * Only where the fused step kernel exists (LP2D_EUNSUPPORTED otherwise); calls that cannot take the fused path on
 * such a plan fail with LP2D_EUNSUPPORTED instead of silently using another layout.  lp2d_transpose converts. */
LP2D_API int lp2d_plan_set_spectrum_layout(lp2d_plan* plan, int transposed);
LP2D_API int lp2d_plan_get_spectrum_layout(lp2d_plan* plan);
/* out[b][j][i] = in[b][i][j] for `batch` planes of n x n complex samples (in != out). */
LP2D_API int lp2d_transpose(int dtype, const void* in, void* out, int n, long long batch, void* stream);

/* Tuning aid (builds with LP2D_BUILD_DEFS="-DLP2D_Q4_CLOCKS=1" only): with LP2D_Q4_TIMING=1 in the environment when the plan is created, the persistent step kernel sums the
 * clock ticks its transform warps spend per phase; host16 receives [row items: wait, tile -> registers, factor,
 * transform + stores, item count, 0, 0, 0 | column items: the same].  LP2D_EINVAL when not enabled. */
LP2D_API int lp2d_q4_timing(lp2d_plan* plan, unsigned long long* host16, int reset);

/* How many kernels this library has launched in this process (for bench accounting). */
LP2D_API long long lp2d_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LP2D_H */
