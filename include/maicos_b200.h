/*
 * maicos_b200 -- C ABI of the B200-native MAICoS hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.
 * The Python host layer (maicos_b200/_cabi.py) binds exactly these symbols with
 * ctypes; a MAICoS maintainer would bind the same symbols from
 * src/maicos/lib/math.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MB200_E* code on failure;
 *     mb200_last_error() returns the message of the last failure on this thread.
 *   - "host" pointers are ordinary (pageable or pinned) CPU memory; the library does
 *     the host<->device copies.  "dev" pointers are CUDA device pointers on the
 *     context's device (e.g. torch.Tensor.data_ptr()).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     MB200_ENODEVICE.
 *   - outputs are written in the layout of the reference (row-major, float64 sums,
 *     int64 counts).
 *
 * Reference file:line citations are relative to /root/reference (MAICoS v0.11.2).
 */
#ifndef MAICOS_B200_H
#define MAICOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_OK 0
#define MB200_EINVAL (-1)
#define MB200_ENODEVICE (-2)
#define MB200_ECUDA (-3)
#define MB200_ENOMEM (-4)
#define MB200_EBUSY (-5) /* no free slot for a deferred batch (mb200_saxs_frames_submit) */

typedef struct mb200_ctx mb200_ctx;

/* ---- library / context -------------------------------------------------------- */

/* ABI version of this header (bumped on any signature change). */
int mb200_abi_version(void);
/* Message of the last error raised on the calling thread ("" if none). */
const char *mb200_last_error(void);
/* Number of visible CUDA devices (0 when there is no driver / no GPU). */
int mb200_device_count(void);
/* One context per (process, device): owns a stream, workspaces and the q-plan cache. */
int mb200_ctx_create(int device, mb200_ctx **out);
void mb200_ctx_destroy(mb200_ctx *ctx);
/* Block until all work queued on the context's stream has finished. */
int mb200_ctx_synchronize(mb200_ctx *ctx);
/* The context's cudaStream_t as an integer (for torch.cuda.ExternalStream / events). */
uint64_t mb200_ctx_stream(mb200_ctx *ctx);
/* Statistics of the last S(q) call on this context (for bench.py / DESIGN.md):
 *   out[0] = kernels launched, out[1] = n_valid q, out[2] = CTAs of the contraction kernel,
 *   out[3] = DMMA tile-steps issued (8x8x4 MMAs), out[4] = atom*q evaluations,
 *   out[5] = last contraction-kernel time in ns (0 unless timing enabled); after mb200_profile_hist: time of
 *            the histogram kernel of that call,
 *   out[6] = k-splits, out[7] = warp tiles in the plan. */
int mb200_ctx_stats(mb200_ctx *ctx, int64_t out[8]);
/* Enable (1) / disable (0) CUDA-event timing of the contraction kernel inside calls. */
int mb200_ctx_set_timing(mb200_ctx *ctx, int enable);
/* Arithmetic of the S(q) contraction (all path-1 entry points):
 *   1  (default) split precision on the tcgen05 tensor cores: the FP64 phase factors are cut into five
 *      balanced base-256 digits (40-bit fixed point), the digit products are accumulated EXACTLY in int32
 *      TMEM accumulators and recombined in FP64.  Measured against mode 0 on every nonzero cell of the raw
 *      cube, 1.5k .. 1M atoms: max relative error on S 4.4e-9, median 6e-12 (tools/tc_cell_error.py);
 *      binned S(q) agrees to 3e-13.  Stated tolerance of the path: 1e-6.
 *   0  FP64 phase tables + FP64 DMMA (mma.sync m8n8k4) accumulation: ~1e-12 against the reference kernel.
 * Plans that cannot be tiled for the tensor cores (a reciprocal cube edge > 255) and calls with non-finite
 * weights use mode 0 automatically.  Weights of mb200_structure_factor[_block] are rescaled by an exact
 * power of two so that any finite weights are admissible in mode 1.
 *   2  as 1, but the generated operand always goes through shared memory (mode 1 writes it straight into tensor
 *      memory when the l tiles of the plan have at most 32 values); same results bit for bit, for A/B measurements. */
int mb200_ctx_set_precision(mb200_ctx *ctx, int mode);
/* Peak probes of the pipes the roofline fractions of bench.py are quoted against, measured in the calling process
 * (a few milliseconds each, best of three):
 *   out[0] = tcgen05 kind::i8 dense rate in TOP/s (M = 128, N = 256, K = 32 from shared memory on every SM),
 *   out[1] = FP64 DMMA rate in TFLOP/s (mma.sync m8n8k4), out[2] = device-to-device copy rate in GB/s
 *   (read + write bytes of a 512 MiB copy), out[3] = scalar FP64 FMA rate in TFLOP/s. */
int mb200_probe_peaks(mb200_ctx *ctx, double out[4]);
/* Total kernels launched by this context since creation. */
int64_t mb200_ctx_launch_count(mb200_ctx *ctx);
/* Announce the host buffer the NEXT mb200_saxs_frames call will read its positions from: the frames are copied to
 * the device on a separate copy stream while the current call computes (host staging -> PCIe -> HBM overlap across
 * calls).  The next call recognises the pointer and skips its own copy.  May be called from another thread than the
 * one inside a compute call; at most four announcements are kept, and a slot a running call still reads is never
 * overwritten.  The buffer must stay untouched until that call. */
int mb200_prefetch_frames(mb200_ctx *ctx, const void *host_frames, size_t bytes);
/* The same for frames that several analyses of one frame loop read (AnalysisCollection, src/maicos/core/base.py:602-718:
 * every instance sees the same untransformed frame): announcements with equal (host_frames, bytes, tag), tag != 0, share
 * ONE host -> device copy, and the slot serves as many compute calls as announced it.  The caller's tag must identify the
 * content (e.g. run number and first frame of the batch): a buffer refilled with other frames needs another tag.
 * mb200_ctx_prefetch_stats: out[0] = announced copies issued since the context was created, out[1] = copies saved by
 * sharing, out[2] = announcements not copied ahead because every slot was being read (the call copies by itself). */
int mb200_prefetch_frames_shared(mb200_ctx *ctx, const void *host_frames, size_t bytes, uint64_t tag);
int mb200_ctx_prefetch_stats(mb200_ctx *ctx, int64_t out[3]);
/* One frame of a GROMACS .xtc file (compressed coordinates), host only.  `data` / `size`: the (memory-mapped) file, `offset`:
 * where the frame starts.  Outputs (each may be NULL): atom count, step, time (ps), box (3 x 3, nm), positions float32
 * [natoms][3] multiplied by `scale` (10: nm -> Angstrom) -- NULL positions only scans the header -- and the offset of the
 * next frame.  Replaces the XTC reader MDAnalysis brings (xdrfile, third party) on the trajectory-ingest side of the path. */
int mb200_xtc_frame(const void *data, size_t size, size_t offset, int64_t *natoms, int32_t *step, float *time,
                    float box[9], float *positions, double scale, size_t *next_offset);
/* Page-locked host memory for frame staging (full-rate asynchronous H2D / D2H).  Any host pointer
 * is accepted by the compute calls; pinned ones avoid the driver's bounce buffer. */
int mb200_host_alloc(size_t bytes, void **out);
void mb200_host_free(void *p);

/* ---- path 1: reciprocal-space structure factor -------------------------------- */

/* maxn[i] = (int)ceil(qmax / (float)(2*pi/dimensions[i]))
 * replaces: src/maicos/lib/_cmath.pyx:93-96 (NB the float32 cast of q_factor). */
int mb200_sfactor_maxn(const double dimensions[3], double qmax, int32_t maxn[3]);

/* Diagnostic, host only (no device needed): builds the q-plan of a box / q window -- the accepted Miller vectors with the
 * reference's acceptance test (_cmath.pyx:93-115), their |q| bins, the DMMA and tensor-core tilings -- and describes it:
 *   out[0] = accepted q vectors, out[1..3] = maxn, out[4] = tensor-core tiling possible, out[5] = row tiles (128 (h,k)
 *   rows), out[6] = l tiles, out[7] = l values per tile, out[8] = populated (row tile, l tile) blocks = contraction items
 *   per atom chunk, out[9] = DMMA warp tiles, out[10..15] = FNV-1a fingerprints of the plan's index tables.
 * n_bins = 0: entries in C order of the cube (mb200_structure_factor), > 0: sorted by numpy |q| bin (the *_frames calls). */
int mb200_plan_describe(const double dimensions[3], double qmin, double qmax, double thetamin, double thetamax,
                        int32_t n_bins, int64_t out[16]);

/* Build the q-plans of the frames of the NEXT *_frames call ahead of it (host work only, no device call): an NPT trajectory
 * has a new cell, hence a new plan (about a millisecond of host time at maxn = 32), in every frame.  Called from the thread
 * that stages the frames while the previous batch computes; the plans of cells that are neither cached nor built already
 * are constructed side by side on the host threads and picked up by the compute call.
 *   boxes : host float64 [n_frames][3] (Saxs: its float32 cell lengths converted to double, as the call does)
 *   n_bins, thetamin / thetamax, cm_params / n_groups : those of the call (DiporderStructureFactor: 0, pi, NULL, 0) */
int mb200_plans_ahead(mb200_ctx *ctx, const double *boxes, int64_t n_frames, double qmin, double qmax, double thetamin,
                      double thetamax, int32_t n_bins, const double *cm_params, int32_t n_groups);

/* Drop-in for `_cmath.structure_factor` (src/maicos/lib/_cmath.pyx:21-126, re-exported
 * at src/maicos/lib/math.py:15; callers saxs.py:197-205, diporderstructurefactor.py:121-129).
 *   positions : host float64, element (j, d) at positions[j*stride_atom + d*stride_xyz]
 *   weights   : host float64 [n_atoms]
 *   scattering_vectors, structure_factors : host float64 [maxn0*maxn1*maxn2], C order,
 *               exactly 0.0 in rejected cells (callers use S != 0 as the validity mask).
 * Acceptance test per cell is the reference's: h+k+l != 0, qmin <= |q| <= qmax,
 * thetamin <= acos(qz/|q|) <= thetamax (radians, inclusive).  */
int mb200_structure_factor(mb200_ctx *ctx, const double *positions, int64_t stride_atom,
                           int64_t stride_xyz, int64_t n_atoms, const double dimensions[3],
                           double qmin, double qmax, double thetamin, double thetamax,
                           const double *weights, double *scattering_vectors,
                           double *structure_factors);

/* Same contraction, restricted to a block of the q-plan ("q-vector blocks as the second
 * axis for single huge frames"): only CTA tile groups g with g % n_blocks == block are
 * computed; all other cells are returned as 0, so summing the outputs of all blocks
 * (e.g. an NCCL allreduce over ranks) reproduces mb200_structure_factor. */
int mb200_structure_factor_block(mb200_ctx *ctx, const double *positions, int64_t stride_atom,
                                 int64_t stride_xyz, int64_t n_atoms,
                                 const double dimensions[3], double qmin, double qmax,
                                 double thetamin, double thetamax, const double *weights,
                                 int32_t block, int32_t n_blocks, double *scattering_vectors,
                                 double *structure_factors);

/* Fused Saxs frame body (replaces the per-group loop of Saxs._single_frame,
 * src/maicos/modules/saxs.py:191-235, bin_spectrum=True): for every frame f and element
 * group g: wrap positions to [-L/2, L/2] in float32 (saxs.py:193-195), S over the valid
 * q set of that frame's box, I = f_g(|q|)^2 * S (math.py:623-702), then the three |q|
 * histograms (saxs.py:222-233) over entries with S != 0.
 *   positions   : float32 [n_frames][n_atoms_total][3] (host, or device if positions_on_device)
 *   boxes       : host float32 [n_frames][3] box lengths (ts.dimensions[:3])
 *   group_index : host int32, concatenated atom indices of the groups
 *   group_ptr   : host int64 [n_groups+1] offsets into group_index
 *   cm_params   : host float64 [n_groups][2][10] = up to two (multiplicity, a1..a4, b1..b4, c)
 *                 Cromer-Mann sets per group (second set for united atoms CHn/NHn, else mult 0)
 *   pack_atoms  : apply AtomGroup.wrap(compound="atoms") (x -= floor(x/L)*L, float32) first
 *   q-plan block selection as in mb200_structure_factor_block (use block=0, n_blocks=1).
 * outputs (host):
 *   s_binned, i_binned : float64 [n_frames][n_groups][n_bins]
 *   bincount           : int64   [n_frames][n_groups][n_bins]  (entries with S != 0)  */
int mb200_saxs_frames(mb200_ctx *ctx, const float *positions, int positions_on_device,
                      int64_t n_frames, int64_t n_atoms_total, const float *boxes,
                      const int32_t *group_index, const int64_t *group_ptr, int32_t n_groups,
                      const double *cm_params, double qmin, double qmax, double thetamin,
                      double thetamax, int32_t n_bins, int pack_atoms, int32_t block,
                      int32_t n_blocks, double *s_binned, double *i_binned, int64_t *bincount);

/* The same with the element groups uploaded ONCE (a selection of a million atoms is 4 MB of indices that the call above
 * validates, fingerprints and -- when they changed -- copies with every batch): create the handle in _prepare, pass it
 * with every batch, destroy it at the end of the run. */
typedef struct mb200_groups mb200_groups;
int mb200_groups_create(mb200_ctx *ctx, const int32_t *group_index, const int64_t *group_ptr, int32_t n_groups,
                        int64_t n_atoms_total, mb200_groups **out);
void mb200_groups_destroy(mb200_ctx *ctx, mb200_groups *groups);
int mb200_saxs_frames_groups(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                             const float *boxes, const mb200_groups *groups, const double *cm_params, double qmin,
                             double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                             int32_t block, int32_t n_blocks, double *s_binned, double *i_binned,
                             int64_t *bincount);

/* Deferred form of mb200_saxs_frames_groups: _submit prepares and queues the batch and returns a ticket (0 or 1) while
 * the kernels run -- or wait behind the previous batch -- so that the host side of the next batch (plan look-up,
 * descriptors, launches) overlaps with the device work of this one; _collect waits for the batch and hands out its results
 * (same layout as mb200_saxs_frames_groups).  Up to four batches of a context may wait at a time (MB200_EBUSY beyond that:
 * collect one, or use the synchronous call).  `boxes` is read before _submit returns; page-locked host `positions` that
 * were not announced with mb200_prefetch_frames are copied by the stream when their turn comes, so they must stay
 * untouched until _collect. */
int mb200_saxs_frames_submit(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                             const float *boxes, const mb200_groups *groups, const double *cm_params, double qmin,
                             double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                             int32_t block, int32_t n_blocks, int32_t *ticket);
int mb200_saxs_frames_collect(mb200_ctx *ctx, int32_t ticket, double *s_binned, double *i_binned, int64_t *bincount);

/* Saxs with bin_spectrum=False (src/maicos/modules/saxs.py:188-189, 238-239: the raw cubes of all element groups are
 * summed per frame) for a batch of frames of ONE cell: per frame and group the structure factors of the accepted q
 * vectors as a compact list in C order of the reciprocal cube -- one tables launch and one contraction launch per batch
 * instead of one mb200_structure_factor call (dense cube back, host scatter) per frame and group.
 *   box      : host float32 [3], the cell of every frame (the class refuses a changing cell, saxs.py:182-186)
 *   n_valid  : number of accepted q vectors = mb200_plan_describe(..., n_bins = 0) out[0] (checked)
 * outputs (host): cells int32 [n_valid] (h * maxn1 * maxn2 + k * maxn2 + l) and scattering_vectors float64 [n_valid]
 *   (either may be NULL), s_values float64 [n_frames][n_groups][n_valid]. */
int mb200_saxs_frames_cells(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                            const float box[3], const mb200_groups *groups, double qmin, double qmax,
                            double thetamin, double thetamax, int pack_atoms, int64_t n_valid,
                            int32_t *cells, double *scattering_vectors, double *s_values);

/* Fused DiporderStructureFactor frame body (replaces
 * src/maicos/modules/diporderstructurefactor.py:105-150): three weight sets on identical
 * positions share one set of phase tables; per frame and component S is binned by |q|.
 *   positions : float64 [n_frames][n][3] (compound centres), host or device (inputs_on_device,
 *               e.g. the output of mb200_compound_prologue)
 *   weights   : float64 [n_frames][3][n], same residence as positions
 *   boxes     : host float64 [n_frames][3]
 * outputs (host): s_binned float64 [n_frames][3][n_bins] (sum of S per bin),
 *                 bincount int64 [n_frames][3][n_bins] (entries with S != 0).  */
int mb200_dipsf_frames(mb200_ctx *ctx, const double *positions, const double *weights,
                       int inputs_on_device, int64_t n_frames, int64_t n, const double *boxes,
                       double qmin, double qmax, int32_t n_bins, int32_t block, int32_t n_blocks,
                       double *s_binned, int64_t *bincount);

/* Staging helper (host only): memcpy of `bytes` split over `n_threads` threads (<= 0: 6).  Used to move a frame from
 * the trajectory reader's buffer (MDAnalysis `ts.positions`) into the pinned batch buffer faster than one core copies. */
int mb200_host_copy(void *dst, const void *src, size_t bytes, int n_threads);

/* Frame statistics of a batch, host only (no device needed): running mean, standard error of the mean and sum of a
 * per-frame observable, frame after frame with the reference's own recurrences
 * (src/maicos/core/base.py:460-480, src/maicos/lib/math.py:325-459), bitwise what the per-frame Python code gives.
 *   x_f64 / x_i64 : [n_frames][size], exactly one of them non-NULL;  n0 >= 1 frames are already in the statistics
 *   mean, sem     : float64 [size], updated in place;  sum_f64 / sum_i64 : [size], the one matching x
 *   scalar_pow    : the observable is a float scalar in the Python frame loop, whose `sem ** 2` is libm's pow(sem, 2) */
int mb200_running_stats(int64_t n0, int64_t n_frames, int64_t size, int scalar_pow, const double *x_f64,
                        const int64_t *x_i64, double *mean, double *sem, double *sum_f64, int64_t *sum_i64);

/* ---- per-frame compound prologue ------------------------------------ */

/* Plain device buffers for intermediate results that stay on the GPU between two entry points. */
int mb200_dev_alloc(mb200_ctx *ctx, size_t bytes, void **out);
int mb200_dev_free(mb200_ctx *ctx, void *p);
int mb200_dev_read(mb200_ctx *ctx, const void *dev, void *host, size_t bytes);
int mb200_dev_write(mb200_ctx *ctx, void *dev, const void *host, size_t bytes);

/* Compound centres and dipole-derived weights from raw atom coordinates, for compounds that are
 * contiguous runs of atoms.  Replaces the per-frame host calls
 *   atoms.unwrap / atoms.wrap(compound)        src/maicos/core/base.py:438-448
 *   get_center(ag, bin_method, compound)       src/maicos/lib/util.py:642-680
 *   diporder_weights(ag, compound, ...)        src/maicos/lib/weights.py:131-178
 *   positions      : float32 [n_frames][n_atoms][3] (host, or device if positions_on_device)
 *   boxes          : host float32 [n_frames][3]
 *   compound_ptr   : host int64 [n_c+1], atoms of compound c are compound_ptr[c] .. compound_ptr[c+1]-1
 *   center_weights : host float64 [n_atoms] (masses for "com", |charges| for "coc") or NULL ("cog")
 *   masses, charges: host float64 [n_atoms]; masses also give the wrap reference point (NULL: geometry)
 *   make_whole     : minimum image about the first atom of each compound (float32)
 *   wrap_center    : shift every compound so that its centre of mass lies in the cell
 *   weight_mode    : 0 none, 1 mu.e_pdim (P0), 2 cos_theta = mu.e_pdim/|mu|, 3 cos_theta^2,
 *                    4 all three components of mu/|mu| laid out [n_frames][3][n_c]
 * outputs (DEVICE): centers_dev float64 [n_frames][n_c][3], weights_dev float64 [n_frames][n_c]
 *                   (or [n_frames][3][n_c] for mode 4); feed them to mb200_profile_hist /
 *                   mb200_dipsf_frames with their *_on_device flags. */
int mb200_compound_prologue(mb200_ctx *ctx, const float *positions, int positions_on_device,
                            int64_t n_frames, int64_t n_atoms, const float *boxes,
                            const int64_t *compound_ptr, int64_t n_c, const double *center_weights,
                            const double *masses, const double *charges, int make_whole,
                            int wrap_center, int weight_mode, int pdim, double *centers_dev,
                            double *weights_dev);

/* The same with the topology side (compound_ptr, centre weights, masses, charges: 8 + 24 bytes per atom) uploaded ONCE:
 * an analysis creates the handle in its _prepare and passes it with every batch of frames.  `positions` announced with
 * mb200_prefetch_frames are taken from the prefetch slot (their copy ran under the previous call). */
typedef struct mb200_topology mb200_topology;
int mb200_topology_create(mb200_ctx *ctx, const int64_t *compound_ptr, int64_t n_c, int64_t n_atoms,
                          const double *center_weights, const double *masses, const double *charges,
                          mb200_topology **out);
void mb200_topology_destroy(mb200_ctx *ctx, mb200_topology *topo);
/*   compound_ptr NULL (n_c == n_atoms): every atom is its own compound (velocity / temperature weights per atom).
 *   uv_mode / uv_dim (weight modes 1..3): the unit vector the dipole is projected on --
 *     0 Cartesian e_pdim                                    src/maicos/lib/util.py:684-706
 *     1 radial direction of a cylinder with axis uv_dim     src/maicos/lib/util.py:710-756  (DiporderCylinder, pdim "r")
 *     2 radial direction of a sphere                        src/maicos/lib/util.py:760-786  (DiporderSphere)
 *   both about the box centre dimensions[:3] / 2, from the compound's binned centre (bin_method). */
int mb200_compound_prologue_topo(mb200_ctx *ctx, const float *positions, int positions_on_device,
                                 int64_t n_frames, const float *boxes, const mb200_topology *topo,
                                 int make_whole, int wrap_center, int weight_mode, int pdim,
                                 int uv_mode, int uv_dim, double *centers_dev, double *weights_dev);
/* The same with the topology's atoms being the window [first_atom, first_atom + n_atoms) of whole frames of n_total
 * atoms (see mb200_profile_hist_window). */
int mb200_compound_prologue_window(mb200_ctx *ctx, const float *positions, int positions_on_device,
                                   int64_t n_frames, int64_t n_total, int64_t first_atom, const float *boxes,
                                   const mb200_topology *topo, int make_whole, int wrap_center, int weight_mode,
                                   int pdim, int uv_mode, int uv_dim, double *centers_dev, double *weights_dev);

/* Velocity-derived weights from the raw velocities of a batch of frames; replaces the per-frame host calls
 *   velocity_weights(ag, grouping, vdim)     src/maicos/lib/weights.py:195-225
 *   temperature_weights(ag, grouping)        src/maicos/lib/weights.py:101-127
 *   velocities : float32 [n_frames][n_atoms][3] (host, or device if velocities_on_device)
 *   mode       : 0 v[:, vdim] per atom, 1 mass-weighted mean of v[:, vdim] per compound, 2 kinetic temperature per atom
 * output (DEVICE): weights_dev float64 [n_frames][n_c], for mb200_profile_hist(weights_on_device=1, weights_per_frame=1). */
int mb200_velocity_weights(mb200_ctx *ctx, const float *velocities, int velocities_on_device, int64_t n_frames,
                           const mb200_topology *topo, int mode, int vdim, double *weights_dev);

/* Per-atom quantities of the "virtual cut" estimator of the dielectric profile modules:
 *   testpos_i = centre of |charge| of atom i's compound along the profile coordinate
 *               (planar: x[dim], dielectricplanar.py:183-185; cylinder: the radial coordinate itself is averaged,
 *               dielectriccylinder.py:167-173), compounds = contiguous runs of atoms;
 *   w_ij      = q_i * (n_j - np.digitize(x_i[cut_dims[j]], arange(n_j)[1:] * step_j))   for cut direction j.
 * Binning testpos with weights w_j through mb200_profile_hist(MB200_GEOM_SCALAR | MB200_BIN_DIGITIZE) gives
 * n_j * mean_c(cumsum_c(curQ[c, :])) of dielectricplanar.py:203-213 / dielectriccylinder.py:180-189 without the
 * (n_j x n_bins) table.
 *   positions : float32 [n_frames][n_atoms][3] (host, or device if positions_on_device)
 *   geometry  : 0 planar, 1 cylinder (center float64 [n_frames][3] required)
 *   cut_bins  : host int32 [n_frames][n_cut], cut_step host float64 [n_frames][n_cut] (= L / n_j)
 * outputs (DEVICE): coord_dev float64 [n_frames][n_atoms], weights_dev float64 [n_frames][n_cut][n_atoms]. */
int mb200_dielectric_prepare(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                             int64_t n_atoms, const int64_t *compound_ptr, int64_t n_c, const double *charges,
                             int geometry, int dim, const double *center, int n_cut, const int32_t *cut_dims,
                             const int32_t *cut_bins, const double *cut_step, double *coord_dev,
                             double *weights_dev);

/* ---- path 2: weighted slab / shell histograms --------------------------------- */

#define MB200_GEOM_PLANAR 0   /* np.histogram(pos[:, dim])            core/planar.py:235-243   */
#define MB200_GEOM_CYLINDER 1 /* transform_cylinder + np.histogram2d  core/cylinder.py:229-243 */
#define MB200_GEOM_SPHERE 2   /* transform_sphere + np.histogram      core/sphere.py:215-223   */
#define MB200_GEOM_SCALAR 3   /* `positions` is the binned coordinate itself, [n_frames][n] (f32 or f64) */
/* or-ed into `geometry`: np.digitize(x, edges[1:-1]) instead of np.histogram -- nothing is dropped, values
 * left / right of the range (and NaN) land in the first / last bin; the cylinder has no axial window.
 * Replaces the digitize + bincount pairs of the dielectric modules:
 *   modules/dielectricplanar.py:170-176, 203-213, modules/dielectriccylinder.py:134-151, 180-190,
 *   modules/dielectricsphere.py:113-133. */
#define MB200_BIN_DIGITIZE 0x10

/* Kernel used for WEIGHTED histograms with finite weights (results are bitwise identical, all accumulate the same
 * exact fixed-point integers):  0 (default) four non-returning shared-memory atomics per element on 18-bit chunks of
 * the fixed-point weight, folded into wide per-bin records every 16384 elements;  1 three returning shared-memory
 * atomics per element with carries (the round-1 kernel);  2 counting-sort tiles, one returning atomic per element (the
 * round-1 default).  1 and 2 are kept for A/B measurements. */
int mb200_ctx_set_hist_mode(mb200_ctx *ctx, int mode);

/* Fused weighted + count histogram of ProfileBase._single_frame (core/base.py:815-818),
 * batched over frames.  numpy semantics (numpy/lib/_histograms_impl.py:816-873, 1037-1069):
 * x is promoted to float64, values outside [edges[0], edges[n_bins]] are dropped, the right
 * edge is inclusive, the bin is the one with edges[i] <= x < edges[i+1].
 *   positions : float32 or float64 [n_frames][n][3] (host or device)
 *   weights   : float64 [n] (weights_per_frame=0) or [n_frames][n] (=1), or NULL
 *   edges     : host float64 [n_frames][n_bins+1]  (np.linspace of the frame's range)
 *   center    : host float64 [n_frames][3] box centre (cylinder / sphere), may be NULL for planar
 *   zrange    : host float64 [n_frames][2] axial window (cylinder only), else NULL
 * outputs (host): hist_w float64 [n_frames][n_bins] (NULL-able when weights NULL),
 *                 hist_c int64 [n_frames][n_bins].  */
int mb200_profile_hist(mb200_ctx *ctx, int geometry, const void *positions, int pos_is_f32,
                       int positions_on_device, int64_t n_frames, int64_t n, int dim,
                       const double *weights, int weights_per_frame, int weights_on_device,
                       const double *edges, int32_t n_bins, const double *center,
                       const double *zrange, double *hist_w, int64_t *hist_c);
/* The same for a group that is a WINDOW of whole frames: `positions` holds [n_frames][n_total][3] (or [n_frames][n_total]
 * for MB200_GEOM_SCALAR) and the group is atoms [first_atom, first_atom + n) of every frame -- a contiguous selection (the
 * water of a slab system) is binned where the trajectory holds the frames, without a gather on the host, and frames
 * announced with mb200_prefetch_frames[_shared] as whole frames serve every analysis of a collection.  Weights are
 * per group atom ([n] or [n_frames][n]) as before.
 *   selection  : NULL, or one group of n atom indices into the whole frames (mb200_groups_create with n_groups = 1 and
 *                n_atoms_total = n_total; first_atom = 0): ANY selection ("type OW") is gathered on the device.
 *   pack_boxes : NULL, or host float32 [n_frames][3] cell lengths -- the float32 positions are first wrapped into the
 *                primary cell on the device, x -= floor(x / L) * L in float32, which is what the reference's frame loop does
 *                on the host for every frame with pack=True and atoms grouping (src/maicos/core/base.py:446-448,
 *                `universe.atoms.wrap(compound="atoms")`). */
int mb200_profile_hist_window(mb200_ctx *ctx, int geometry, const void *positions, int pos_is_f32,
                              int positions_on_device, int64_t n_frames, int64_t n_total, int64_t first_atom,
                              int64_t n, const mb200_groups *selection, const float *pack_boxes, int dim,
                              const double *weights, int weights_per_frame,
                              int weights_on_device, const double *edges, int32_t n_bins, const double *center,
                              const double *zrange, double *hist_w, int64_t *hist_c);

#ifdef __cplusplus
}
#endif
#endif /* MAICOS_B200_H */
