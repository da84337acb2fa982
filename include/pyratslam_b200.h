/*
 * pyratslam_b200.h -- C ABI of the B200-native pose-cell / view-template hot path.
 *
 * This is the drop-in boundary.  The reference (bjkomer/pyratslam) has no FFI of
 * its own: its operator layer is the Python class `Convolution`
 * (ratslam/convolution.py:10-697, OpenCL via pyopencl), driven by
 * `PoseCellNetwork` (ratslam/posecell_network.py:22-353), and the numpy loop in
 * `ViewTemplates.match` (ratslam/view_templates.py:63-75).  Each entry point below
 * names the reference code it replaces.  Plain pointers and sizes only; `stream`
 * is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *
 * Every function returns 0 on success or a negative PRS_E_* code;
 * `prs_last_error()` gives the message of the last failure on the calling thread.
 * There is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * Pose-cell state layout in HBM ("theta-major"):  state[b][th][x][y], y fastest,
 * element type float (dtype 0) or double (dtype 1).  The reference's numpy layout
 * is [x][y][th]; `prs_pc_import_xyt` / `prs_pc_export_xyt` convert.
 */
#ifndef PYRATSLAM_B200_H
#define PYRATSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PRS_API __attribute__((visibility("default")))
#else
#define PRS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PRS_OK 0
#define PRS_E_INVALID (-1) /* bad argument (shape, dtype, null pointer) */
#define PRS_E_CUDA (-2)    /* CUDA runtime error, see prs_last_error() */
#define PRS_E_NODEVICE (-3)
#define PRS_E_LUT_KEY (-4) /* prs_pc_update_host: the reference raises KeyError for this odometry */
#define PRS_E_RADIUS (-5)  /* prs_pc_update_host: the translation does not fit the grid */

#define PRS_F32 0
#define PRS_F64 1
#define PRS_U8 2 /* view-template libraries only */

/* per-network error bits written by the step kernels into `err[b]` */
#define PRS_ERR_LUT_KEY 1 /* fractional x offset exactly +0.5: the reference raises KeyError
                             (posecell_network.py:249) */
#define PRS_ERR_RADIUS 2  /* 3+ceil|vtrans/0.2| > min(X,Y): the reference convolves unwritten
                             memory (convolution.py:661-675) */
#define PRS_ERR_THETA 4   /* |floor(vrot/(2pi/Th)+.5)| beyond the filter table */

#define PRS_OG_RANGE 8    /* theta-origin filters are tabulated for og in [-8, 8] */

#define PRS_VT_MODE_REF 0      /* 15 windowed row offsets, view_templates.py:16-28 */
#define PRS_VT_MODE_CIRCULAR 1 /* all 32 cyclic row shifts (extension; BASELINE config 5) */

PRS_API const char* prs_last_error(void);
PRS_API int prs_version(void);
/* number of CUDA devices visible, or a negative error code */
PRS_API int prs_device_count(void);

/* ------------------------------------------------------------------ pose cells */

/* Host-side tables, all float64, computed by the caller exactly as the reference
 * computes them (numpy/scipy on the host) so that no device libm result can change
 * a filter.  Copied to the device by prs_pc_create. */
typedef struct prs_pc_config {
  int X, Y, Th;          /* PoseCellNetwork(shape), posecell_network.py:24-26 */
  int B;                 /* number of independent networks (1 for the drop-in class) */
  int dtype;             /* PRS_F32 or PRS_F64 */
  double vtrans_scale;   /* pc_vtrans_scale = 0.2, posecell_network.py:35 */
  double vrot_scale;     /* pc_vrot_scale = 2*pi/Th, posecell_network.py:36 */
  const double* ge;      /* [7]  exp(-d^2/(2 sigma_e^2)) : kernel_3d == aE ge(x)ge(x)ge - aI gi(x)gi(x)gi */
  const double* gi;      /* [7]  exp(-d^2/(2 sigma_i^2))   (posecell_network.py:97-113) */
  double aE, aI;         /* amplitudes including the 1/|sum| normalisation */
  const double* f2d;     /* [2][7][7] LUT filters F0 (origin 0,0) and F-1 (origin -1,-1),
                            posecell_network.py:50-59,210-222 */
  const double* f1d;     /* [2*PRS_OG_RANGE+1][7] theta filters for og=-8..8, posecell_network.py:224-235 */
  const double* cos_th;  /* [Th] cos((k-mid)*vrot_scale), posecell_network.py:260-261 */
  const double* sin_th;  /* [Th] sin((k-mid)*vrot_scale) */
} prs_pc_config;

typedef struct prs_pc_plan* prs_pc_handle;

/* replaces Convolution.__init__/set_params/set_text/build_program (convolution.py:11-35,37-91,94-398) */
PRS_API int prs_pc_create(const prs_pc_config* cfg, prs_pc_handle* out);
PRS_API int prs_pc_destroy(prs_pc_handle h);
/* bytes of one full state tensor [B][Th][X][Y] */
PRS_API size_t prs_pc_state_bytes(prs_pc_handle h);
/* which kernel family a step uses: 0 = generic multi-kernel path, 1 = fused SMEM-resident kernel (one CTA per
 * network), 2 = tiled large-grid kernels, 3 = one network per thread-block cluster */
#define PRS_PATH_AUTO (-1)
#define PRS_PATH_GENERIC 0
#define PRS_PATH_RESIDENT 1
#define PRS_PATH_TILED 2
#define PRS_PATH_CLUSTER 3
#define PRS_PATH_PAIR 4 /* fused SMEM-resident kernel, one network per 2-CTA cluster (float32 and float64) */
PRS_API int prs_pc_path(prs_pc_handle h);
/* force the generic path (1) or let the plan choose (0); for tests and profiling */
PRS_API int prs_pc_force_generic(prs_pc_handle h, int on);
/* choose the kernel family explicitly (PRS_PATH_*, PRS_PATH_AUTO = the plan's own choice); PRS_E_INVALID if the
 * plan's shape / dtype is not supported by that family; for tests and profiling */
PRS_API int prs_pc_set_path(prs_pc_handle h, int path);
/* Per-plan options.  PRS_OPT_TILED_TMA: the tiled (large-grid) family runs its 7x7 + theta stages as ONE kernel fed by
 * TMA tensor copies from a padded, origin-aligned intermediate (value 1) or as two kernels (value 0, the default: the
 * fused kernel is parity-equal but measured slower on B200). */
#define PRS_OPT_TILED_TMA 0
/* PRS_OPT_TILED_DOG: the tiled family runs the theta, y and x passes of the DoG as ONE kernel (shared-memory ring fed by
 * cp.async, no (E, I) intermediate in L2) instead of two; parity-equal, measured slower, off by default. */
#define PRS_OPT_TILED_DOG 1
/* PRS_OPT_ACTIVE_SET: sparsity-aware update (csrc/posecell_active.cu).  The attractor dynamics of
 * posecell_network.py:326-353 keep the activity in a compact packet (tens of non-zero cells out of thousands); with this
 * option an update does the reference's arithmetic only where the result can be non-zero and writes the state in place,
 * instead of the dense kernels whose cost does not depend on the state.  Exact for any state: a network whose non-zero
 * cells do not fit the kernel's lists / shared memory, or whose global inhibition is negative, is updated by the dense
 * kernels in the same call.  0 = off (default), 1 = the state is scanned for its non-zero cells every update,
 * 2 = the list of non-zero cells an update leaves behind is the next update's input (no pass over the state at all);
 * every library call that writes the state invalidates the lists, a caller that writes the state by other means must
 * call prs_pc_invalidate_active.  Dimensions up to 256 per axis. */
#define PRS_OPT_ACTIVE_SET 2
PRS_API int prs_pc_set_option(prs_pc_handle h, int option, int value);
/* the state was written by something other than this library's calls: forget the active lists (PRS_OPT_ACTIVE_SET = 2) */
PRS_API int prs_pc_invalidate_active(prs_pc_handle h, void* stream);

/* One PoseCellNetwork.update() for all B networks (posecell_network.py:326-353):
 *   state  : device, [B][Th][X][Y] of the plan's dtype, updated in place
 *   odom   : device, double [B][2] = (vtrans, vrot) as passed to update()
 *   gi     : device, [B] of the plan's dtype, global inhibition (posecell_network.py:34)
 *   argmax : device, int64 [B]  flat index x*Y*Th + y*Th + th of the first maximum
 *            (numpy.argmax order, posecell_network.py:317-319)
 *   total  : device, [B] of the plan's dtype, the sum before normalisation (posecell_network.py:343)
 *   err    : device, int32 [B], PRS_ERR_* bits (0 = fine)
 * Calls on one plan must be ordered by the stream(s) they are enqueued on (as for any in-place update of `state`).  On the
 * fused one-CTA-per-network path consecutive calls on a stream overlap on the device -- update n+1 of a network starts as
 * soon as its update n is complete, not when the whole launch n is (per-network sequence numbers, posecell_resident.cu);
 * a launch that was NOT ordered behind its predecessor would wait for it in vain and traps after two seconds.
 */
PRS_API int prs_pc_step(prs_pc_handle h, void* state, const double* odom, const void* gi, long long* argmax,
                void* total, int* err, void* stream);
/* T consecutive updates with odometry odom[T][B][2]; argmax[T][B], total[T][B]; err[B] is OR-ed */
PRS_API int prs_pc_run(prs_pc_handle h, void* state, const double* odom, int T, const void* gi,
               long long* argmax, void* total, int* err, void* stream);
/* Same as prs_pc_step with HOST odometry/results: H2D copy, step, D2H copy, stream sync.
 * odom_host/argmax_host/err_host should be pinned for the copies to be asynchronous. */
PRS_API int prs_pc_step_host(prs_pc_handle h, void* state, const double* odom_host, const void* gi,
                     long long* argmax_host, int* err_host, void* stream);

/* As prs_pc_step_host, but the result comes back as ONE array int32 result_host[B][4] = (x, y, th, err):
 * the arg-max already unravelled the way get_pc_max does (posecell_network.py:318) plus the PRS_ERR_* bits. */
PRS_API int prs_pc_step_host_xyz(prs_pc_handle h, void* state, const double* odom_host, const void* gi, int* result_host,
                         void* stream);
/* The same step WITHOUT the host waiting for it.  The plan keeps two slots of device staging and two copy streams:
 * the odometry of this step is copied in (from pinned odom_host) on a copy stream, the update runs on `stream`, the
 * packed result is copied out (to pinned result_host) on another copy stream.  With one step kept in flight -- two
 * (odom_host, result_host) pairs used alternately -- every step still moves its inputs in and its result out, but
 * those copies and the launch overhead overlap the neighbouring steps' kernels.  *slot_out (0 or 1) names the step for
 * prs_pc_host_result_wait, which blocks until result_host of that step is complete.  At most two steps in flight. */
PRS_API int prs_pc_step_host_xyz_async(prs_pc_handle h, void* state, const double* odom_host, const void* gi,
                               int* result_host, void* stream, int* slot_out);
PRS_API int prs_pc_host_result_wait(prs_pc_handle h, int slot);
/* PoseCellNetwork.update((vtrans, vrot)) of a single-network plan (B == 1) in one call (posecell_network.py:326-353).
 * The odometry is first checked on the host, in numpy's float64 arithmetic, for the cases the reference cannot
 * handle -- PRS_E_LUT_KEY: a fractional x offset of exactly +0.5 (KeyError, posecell_network.py:249); PRS_E_RADIUS:
 * 3 + ceil|vtrans/0.2| > min(X, Y) (convolution.py:661-675 reads unwritten memory) -- and such an update leaves the
 * state untouched.  Otherwise as prs_pc_step_host_xyz with odom_pinned[0..1] = (vtrans, vrot); both buffers pinned. */
PRS_API int prs_pc_update_host(prs_pc_handle h, void* state, double vtrans, double vrot, const void* gi,
                       double* odom_pinned, int* result_pinned, void* stream);

/* PoseCellNetwork.path_integration() alone (posecell_network.py:252-314): per-plane shifted 7x7
 * correlate + clamp, theta correlate + clamp; no attractor dynamics, no normalisation. */
PRS_API int prs_pc_path_integration(prs_pc_handle h, void* state, const double* odom, int* err, void* stream);

/* posecells[loc] += energy for network b, or for every network when b < 0 (posecell_network.py:322-324) */
PRS_API int prs_pc_inject(prs_pc_handle h, void* state, int b, int x, int y, int th, double energy, void* stream);
/* arg-max without an update (get_pc_max, posecell_network.py:317-319) */
PRS_API int prs_pc_argmax(prs_pc_handle h, const void* state, long long* argmax, void* stream);
/* Sparse read-back: the cells of network b with activity > threshold, in the reference's C order -- what the
 * viewers draw (`nonzero(pc > .002)`, simulate.py:61; ratslam_viewer.py:137) -- instead of copying the whole grid.
 *   idx_out : device int32 [max_out] flat indices x*Y*Th + y*Th + th (ascending)
 *   val_out : device [max_out] of the plan's dtype
 *   count_out : device int32, the number of cells above the threshold (may exceed max_out; the rest is dropped)
 *   work : device scratch of prs_pc_active_work_bytes(h) */
PRS_API size_t prs_pc_active_work_bytes(prs_pc_handle h);
PRS_API int prs_pc_active_cells(prs_pc_handle h, const void* state, int b, double threshold, int max_out, int* idx_out,
                        void* val_out, int* count_out, void* work, void* stream);
/* layout conversion between the reference's [B][x][y][th] and theta-major [B][th][x][y] (device to device) */
PRS_API int prs_pc_import_xyt(prs_pc_handle h, void* state, const void* xyt, void* stream);
PRS_API int prs_pc_export_xyt(prs_pc_handle h, const void* state, void* xyt, void* stream);

/* --------------------------------------------------------------- view templates */

/* Sub-sample a camera frame into a template: rows/cols strictly inside (lo, hi) with
 * (idx-lo) % step != 0 -- the boolean mask of view_templates.py:48-57 applied as in :64.
 * frame: device uint8 [im_rows][im_cols]; out: device uint8 [n_rows][n_cols]. */
PRS_API int prs_vt_extract_u8(const uint8_t* frame, int im_rows, int im_cols, int row_lo, int row_hi, int row_step,
                      int col_lo, int col_hi, int col_step, uint8_t* out, int n_rows, int n_cols, void* stream);

/* Sweep a library of n 32x32 templates against one query (ViewTemplates.match line 65 and
 * ViewTemplate.match, view_templates.py:16-28,65):
 *   key_out : device uint64, receives (min_score << 32) | (base_index + argmin), ties -> lowest
 *             index (numpy.argmin, view_templates.py:73).  n == 0 leaves key_out = UINT64_MAX.
 *   scores  : optional device uint32 [n] (u8) / float [n] (f32): per-template min score, or NULL.
 * u8: score = sum((a - b) mod 256) exactly as numpy uint8 arithmetic gives it.
 * f32: score = sum |a - b| in float32; the key holds the IEEE bits of the score. */
PRS_API int prs_vt_sweep_u8(const uint8_t* lib, long long n, const uint8_t* query, int mode, long long base_index,
                    unsigned long long* key_out, uint32_t* scores, void* stream);
PRS_API int prs_vt_sweep_f32(const float* lib, long long n, const float* query, int mode, long long base_index,
                     unsigned long long* key_out, float* scores, void* stream);
/* Bit-sliced ("packed") uint8 library: the layout the fast sweep streams.  1088 bytes per template, stored
 * in groups of 32 (see csrc/view_templates.cu).  Same scores, keys and tie-breaks as prs_vt_sweep_u8.
 *   prs_vt_packed_bytes(n)  : bytes of a packed library with room for n templates (whole groups)
 *   prs_vt_pack_u8          : convert n row-major templates src[n][32][32] into slots first..first+n-1
 *                             (the packed buffer must have been zero-filled once; ViewTemplates' append, :68-71)
 *   prs_vt_unpack_u8        : one template back to row-major (the ViewTemplate.template attribute)
 *   prs_vt_sweep_packed_u8  : query is a row-major device uint8[32][32]; scratch is a device buffer of
 *                             >= 2 KiB, private to the caller.  The query planes go through one constant-memory
 *                             buffer per device: sweeps issued on different streams (or host threads) of a device
 *                             are correct but run one after the other on the device (event-ordered, the host
 *                             does not block). */
PRS_API size_t prs_vt_packed_bytes(long long n);
PRS_API int prs_vt_pack_u8(const uint8_t* src, long long n, void* packed, long long first, void* stream);
PRS_API int prs_vt_unpack_u8(const void* packed, long long index, uint8_t* dst, void* stream);
PRS_API int prs_vt_sweep_packed_u8(const void* packed, long long n, const uint8_t* query, int mode, long long base_index,
                           unsigned long long* key_out, uint32_t* scores, void* scratch, void* stream);
/* Tuning knobs of the sweeps, for profiling (defaults are the measured best): knob 0 = ring depth of the packed
 * reference-mode sweep (slots of 2 KiB per warp fed by TMA bulk copies; 0 = the register-prefetch kernel, else 2, 4
 * or 8; 34 = four slots and the warps of a CTA kept in lock step by a barrier per ring item, 44 = the same with one
 * 640-thread CTA per SM), knob 1 = CTAs per SM that sweep's grid is sized for, knob 2 = ring depth of the float32 sweep (0 = the register
 * kernel; 1..4 = templates in flight per warp, one template per warp; 11..13 = the column-pair kernel -- two columns per
 * lane, two templates per warp -- with 1..3 template pairs in flight per warp), knob 3 = CTAs per SM of that sweep, knob 4 = the
 * circular-mode sweep keeps the warps of a CTA in lock step (1, default) or not (0). */
PRS_API int prs_vt_tune(int knob, int value);
/* Any template shape rows x cols (row-major library) and any max_offset (view_templates.py:14): the
 * reference's windowed match for configurations other than its 32x32 default.  Correctness path. */
PRS_API int prs_vt_sweep_any_u8(const uint8_t* lib, long long n, const uint8_t* query, int rows, int cols, int max_offset,
                        long long base_index, unsigned long long* key_out, uint32_t* scores, void* stream);
PRS_API int prs_vt_sweep_any_f32(const float* lib, long long n, const float* query, int rows, int cols, int max_offset,
                         long long base_index, unsigned long long* key_out, float* scores, void* stream);
/* HOST query in, HOST key out: H2D copy of the 1 KiB query, sweep, D2H of the 8-byte key, sync.
 * lib stays resident on the device; scratch is a device buffer of >= 1024+8 bytes. */
PRS_API int prs_vt_match_host_u8(const uint8_t* lib, long long n, const uint8_t* query_host, int mode,
                         long long base_index, unsigned long long* key_host, void* scratch, void* stream);

/* ------------------------------------------- sharded library: the MIN over ranks, on the device */

/* A very large library is split by contiguous template ranges over the ranks of a job, one process per GPU
 * (BASELINE config 5).  The reference has no counterpart (it is single-process); the semantics to keep are those of
 * `vals.index(min(vals))` / numpy.argmin over the whole list (ratslam/view_templates.py:65-73): minimum score, lowest
 * global index among equals -- i.e. the MIN of the ranks' packed keys -- followed by the strict '>' threshold test.
 *
 * prs_xchg is a small exchange buffer in device memory that the peers map through CUDA IPC (NVLink peer access):
 *   prs_xchg_create   allocate this rank's buffer (world <= 16)
 *   prs_xchg_export   64-byte cudaIpcMemHandle_t of the buffer, to be all-gathered by the host (any transport)
 *   prs_xchg_connect  open the peers' handles, handles = [world][64] in rank order (the own entry is ignored);
 *                     a world of 1 needs neither export nor connect
 *   prs_xchg_set_timeout  bound of the device-side wait for the peers (default 10 s; then status = 1)
 * One kernel per query publishes the rank's key(s) into every peer's buffer with system-scope stores, waits for the
 * peers' keys of the same query, and reduces them; every rank must issue the same sequence of exchange calls. */
typedef struct prs_xchg prs_xchg;
#define PRS_XCHG_HANDLE_BYTES 64
typedef struct prs_shard_result {
  unsigned long long key;   /* MIN over the ranks of (score << 32 | global index); UINT64_MAX: nothing compared */
  int created;              /* 1: the query became template `template_index` (appended on the owning rank) */
  int template_index;       /* what ViewTemplates.match(...).get_index() returns; -1 from prs_vt_shard_exchange */
  int n_total;              /* library size over all ranks after this query */
  int status;               /* 0 = fine, 1 = a peer did not publish within the timeout */
  unsigned long long seq;   /* sequence number of the exchange (1, 2, ...) */
} prs_shard_result;
PRS_API int prs_xchg_create(int world, int rank, prs_xchg** out);
PRS_API int prs_xchg_export(prs_xchg* x, void* handle64);
PRS_API int prs_xchg_connect(prs_xchg* x, const void* handles);
PRS_API int prs_xchg_set_timeout(prs_xchg* x, double seconds);
PRS_API int prs_xchg_destroy(prs_xchg* x);
/* MIN over the ranks of n_keys (<= 64) packed keys: keys_local is this rank's (device), keys_out receives the global
 * minima (device memory or pinned host memory; optional when result is given: result->key is the first key's minimum),
 * result (optional, device or pinned host) the status record. */
PRS_API int prs_vt_shard_exchange(prs_xchg* x, const unsigned long long* keys_local, int n_keys,
                          unsigned long long* keys_out, prs_shard_result* result, void* stream);
/* ViewTemplates.match's decision for a sharded library (view_templates.py:67-73), taken on the device by every rank
 * from the MIN of the ranks' keys: create when the library is empty or score > threshold (strict; the score is the
 * integer sum for PRS_U8, the float32 sum for PRS_F32, compared in double), else match the key's index.  On the
 * owning rank (owner != 0) a created template -- tpl, the query as a row-major 32x32 template on the device -- is
 * stored in slot n_local of lib (bit-sliced for PRS_U8, row-major for PRS_F32; the buffer must have room).
 * result: device or pinned host memory, valid once `stream` has completed. */
PRS_API int prs_vt_shard_decide(prs_xchg* x, const unsigned long long* key_local, double threshold, int dtype,
                        const void* tpl, void* lib, long long n_local, long long n_total, int owner,
                        prs_shard_result* result, void* stream);
/* Host side of a polled exchange: with `result` in pinned host memory, spin until the LAST exchange issued on x has
 * written its record (its sequence number goes out last) -- about 1 us after the kernel, where a stream
 * synchronisation costs a thread wake-up.  PRS_E_CUDA after timeout_s seconds. */
PRS_API int prs_xchg_wait(prs_xchg* x, const prs_shard_result* result_host, double timeout_s);
/* One query of a sharded library in one call: local sweep (PRS_U8: bit-sliced library + scratch >= 2 KiB as for
 * prs_vt_sweep_packed_u8; PRS_F32: row-major library), exchange (decide == 0: result->key is the MIN over the ranks;
 * decide != 0: the create-or-match decision of prs_vt_shard_decide as well), wait for the pinned record.
 * key_dev: device uint64 scratch for the local key. */
PRS_API int prs_vt_shard_query(prs_xchg* x, int dtype, void* lib, long long n_local, const void* query_dev, int mode,
                       long long base_index, unsigned long long* key_dev, void* scratch, int decide, double threshold,
                       long long n_total, int owner, prs_shard_result* result_pinned, void* stream);

/* ------------------------------------------------------------------- one frame */

/* Result of prs_frame_host (32 bytes, written to pinned host memory). */
typedef struct prs_frame_result {
  long long argmax;          /* arg-max pose cell after the update (flat index), or the previous one */
  unsigned long long key;    /* (best score << 32) | best index, UINT64_MAX if the library was empty */
  int created;               /* 1: the frame became template `template_index` */
  int template_index;        /* what ViewTemplates.match(...).get_index() returns */
  int n_templates;           /* library size after this frame */
  int pc_err;                /* PRS_ERR_* bits of the pose-cell update */
} prs_frame_result;

PRS_API size_t prs_frame_scratch_bytes(void);
/* One iteration of the ROS loop (ratslam/ros_simulate.py:98-105,134-137) with a single host sync:
 * [odom_host != NULL: pose-cell update with (vtrans, vrot)] -> frame H2D -> sub-sample -> sweep of the packed
 * uint8 library -> create-or-match on the device (a created template is packed into slot n_templates, so
 * the library must have room for n_templates + 1) -> result D2H.
 *   pc_work : device, 24 bytes (arg-max, total, err of the single network); persists between frames
 *   scratch : device, prs_frame_scratch_bytes()
 *   frame_host / odom_host / result_host : host (pinned for asynchronous copies) */
PRS_API int prs_frame_host(prs_pc_handle pc, void* pc_state, const void* gi, void* pc_work, const double* odom_host,
                   void* vt_packed, int n_templates, unsigned threshold, int mode, const uint8_t* frame_host,
                   int im_rows, int im_cols, int row_lo, int row_hi, int row_step, int col_lo, int col_hi,
                   int col_step, void* scratch, prs_frame_result* result_host, void* stream);

/* The same frame as a replayed CUDA graph: all pointers (pinned host buffers included) are fixed at creation and
 * the library size lives in device memory, so one frame = one graph launch + one synchronisation.
 * The library must keep a free slot: recreate the plan when ViewTemplates grows its buffer.
 * prs_frame_run(plan, moved, stream): moved != 0 -> odom_host holds (vtrans, vrot) and the pose cells update first;
 * `stream` must be a non-default stream for the graph to be captured (otherwise every frame is enqueued eagerly). */
typedef struct prs_frame_plan prs_frame_plan;
PRS_API int prs_frame_create(prs_pc_handle pc, void* pc_state, const void* gi, void* pc_work, void* vt_packed,
                     int n_templates, int capacity, unsigned threshold, int mode, int im_rows, int im_cols,
                     int row_lo, int row_hi, int row_step, int col_lo, int col_hi, int col_step, void* scratch,
                     double* odom_host, uint8_t* frame_host, prs_frame_result* result_host, prs_frame_plan** out);
PRS_API int prs_frame_destroy(prs_frame_plan* plan);
/* Put the library size back into the plan's device-side counter after templates were appended outside the plan
 * (ViewTemplates.match / create between fused frames); synchronises `stream`. */
PRS_API int prs_frame_set_count(prs_frame_plan* plan, int n_templates, void* stream);
PRS_API int prs_frame_run(prs_frame_plan* plan, int moved, void* stream);
/* prs_frame_run without the final synchronisation: the caller waits on `stream` (or an event recorded on it) before
 * reading result_host.  Two plans that share pc / pc_work / scratch / vt_packed (hence the device-side template
 * count) but have their own pinned host buffers can be launched alternately on ONE stream: the device runs the
 * frames back to back while the host prepares the next frame and digests the previous result. */
PRS_API int prs_frame_launch(prs_frame_plan* plan, int moved, void* stream);

/* The whole loop of ros_simulate.py:152-166 over recorded arrays in one call: frame t is staged into the pinned
 * buffers of plans[t % n_plans] and launched while earlier frames are still running; a plan is reused once its frame
 * has finished and its result has been copied to results[].  The plans must share pc / pc_state / pc_work / vt_packed /
 * scratch and differ only in their pinned host buffers; the library needs room for T more templates.
 *   frames : host uint8 [T][im_rows][im_cols]     odom : host double [T][2] = (vtrans, vrot) as given to update()
 *   moved  : host uint8 [T], 0 = the twist was below the 0.001 gate (ros_simulate.py:128): no pose-cell update
 *   results: host prs_frame_result [T]            stream: a non-default stream */
PRS_API int prs_replay_run(prs_frame_plan* const* plans, int n_plans, const uint8_t* frames, const double* odom,
                   const uint8_t* moved, int T, prs_frame_result* results, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYRATSLAM_B200_H */
