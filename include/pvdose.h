/* pvdose.h - C ABI of libpvdose.so: B200-native (sm_100a) kernel-convolution dose path.
 *
 * This is the drop-in boundary for ONE hot path of devhliu/PyVoxelDosimetry.  The reference has
 * no FFI of its own (it is pure Python); each entry point below names the reference interface
 * (file:line under /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer adds to the reference.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.
 *  - All array pointers are DEVICE pointers owned by the caller unless the name starts with
 *    `h_` (host).  Volumes are dense C-order [n0][n1][n2] float32 (the reference's NumPy
 *    arr[x, y, z] layout, z contiguous: examples/kernel_convolution_example.py:17).
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised, unless
 *    stated.
 *  - Every function returns 0 on success or a negative PVD_ERR_* code; pvd_last_error()
 *    returns a thread-local message.  No hidden device allocation: the caller provides the
 *    workspace (pvd_plan_workspace_bytes / pvd_plan_set_workspace).
 */
#ifndef PVDOSE_H
#define PVDOSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVD_VERSION 100

#define PVD_OK 0
#define PVD_ERR_INVALID (-1)     /* bad argument */
#define PVD_ERR_CUDA (-2)        /* CUDA runtime error (message has the cudaError string) */
#define PVD_ERR_STATE (-3)       /* workspace / kernel not set */
#define PVD_ERR_NONFINITE (-4)   /* dose kernel contains NaN/Inf (the reference's own Y90 kernel does:
                                    data/dose_kernels/y90_kernel.py:134-138) */
#define PVD_ERR_UNSUPPORTED (-5) /* size outside what the engine handles */

/* Boundary semantics of the convolution. */
#define PVD_BOUNDARY_REFERENCE 0 /* circular over the activity grid, kernel anchored (cropped / zero padded) at
                                    index (0,0,0): exactly np.fft.ifftn(fftn(a) * fftn(k, a.shape)).real,
                                    core/kernel_convolution.py:71-74 */
#define PVD_BOUNDARY_SAME 1      /* zero boundary, kernel centred at k//2 (generator convention,
                                    data/dose_kernels/y90_kernel.py:34): == reference op on the zero padded
                                    volume, cropped at the kernel centre */

#define PVD_ALGO_AUTO 0
#define PVD_ALGO_FFT 1    /* hand-written Stockham 3-D real FFT convolution */
#define PVD_ALGO_DIRECT 2 /* direct tiled convolution, TMA-staged halo tiles (small kernels) */
#define PVD_ALGO_FFT_UNPIPELINED 3 /* FFT path on the one-tile-per-CTA kernels only: the form the engine falls back to where
                                      the persistent, TMA-pipelined tiles do not fit one SM's shared memory (lengths 864,
                                      1024, 1152); selectable so that validation can cover it at any size */

#define PVD_MAX_T 16 /* activity volumes fused into one execute call */

typedef struct pvd_plan pvd_plan;

typedef struct pvd_plan_info {
    int n[3];      /* input extents */
    int m[3];      /* transform extents */
    int out_lo[3]; /* first output index per axis (in transform coordinates) */
    int out_n[3];  /* output extents */
    int k[3];      /* kernel extents */
    int algo;      /* PVD_ALGO_FFT or PVD_ALGO_DIRECT (AUTO resolved) */
    int passes;    /* kernel launches per execute */
    size_t workspace_bytes;
    double hbm_bytes_per_execute; /* bytes the implementation moves through HBM per execute (T=1, density) */
} pvd_plan_info;

int pvd_version(void);
/* Hash of the CUDA sources this library was built from (stamped by __graft_entry__.build()); measured artefacts that depend
 * on the kernels (ncu DRAM traffic) carry it so that stale ones can be told apart. */
const char* pvd_build_id(void);
const char* pvd_last_error(void);

/* Smallest transform length >= n the engine handles efficiently ({2,3,5,7}-smooth, few radix stages). */
int pvd_good_fft_size(int n);
/* Same for a given axis (0, 1: strided column passes, 2: contiguous row passes - they have different menus of
 * size-specialised kernels). */
int pvd_good_fft_size_axis(int n, int axis);

/* ---- A1: KernelConvolutionCalculator.calculate_dose_rate (core/kernel_convolution.py:48-76) ----
 * Plan for convolving [n0][n1][n2] activity volumes with a [k0][k1][k2] dose voxel kernel. */
int pvd_plan_create(pvd_plan** out, const int n[3], const int k[3], int boundary, int algo);

/* Expert form used by the z-slab decomposition (slab-plus-halo input, interior output):
 * transform extents m[i] (0 = choose), outputs taken from transform indices
 * [out_lo[i], out_lo[i] + out_n[i]).  The kernel is anchored at the origin of the circular
 * domain of extents m, so out_lo = k//2 with m >= n + k - 1 - k//2 gives `same` semantics. */
int pvd_plan_create_ex(pvd_plan** out, const int n[3], const int m[3], const int out_lo[3], const int out_n[3],
                       const int k[3], int algo);

int pvd_plan_get_info(const pvd_plan* plan, pvd_plan_info* info);
int pvd_plan_workspace_bytes(const pvd_plan* plan, size_t* bytes);
/* Workspace must be 256-byte aligned device memory of at least workspace_bytes; it holds the twiddle
 * tables, the complex work buffer and the cached kernel spectrum. */
int pvd_plan_set_workspace(pvd_plan* plan, void* workspace, size_t bytes, void* stream);

/* Upload the dose voxel kernel (dense [k0][k1][k2] float32, device).  Validates finiteness
 * (synchronises `stream` once) and builds the cached spectrum: the reference recomputes
 * fftn(kernel) on every call (core/kernel_convolution.py:73). */
int pvd_plan_set_kernel(pvd_plan* plan, const float* kernel, void* stream);

/* dose = scale * conv( sum_t h_weights[t] * act[t], kernel ) [* rho_ref / max(rho, rho_min)]
 *  - T = 1, h_weights = NULL           : A1 dose rate (core/kernel_convolution.py:71-74)
 *  - T > 1, trapezoid weights          : A2 calculate_absorbed_dose (core/kernel_convolution.py:94-106)
 *                                        evaluated as ONE convolution by linearity
 *  - density != NULL                   : A9 voxel-wise density correction (the `tissue_densities`
 *                                        argument the reference accepts and ignores, core/dose_calculator.py:90);
 *                                        voxels with rho < rho_cut are zeroed (rho_cut <= 0 disables)
 * h_act is a HOST array of T device pointers; h_weights a HOST array of T floats (NULL = all 1). */
int pvd_conv_execute(pvd_plan* plan, const float* const* h_act, const float* h_weights, int T,
                     const float* density, float rho_ref, float rho_min, float rho_cut, float scale, float* dose,
                     void* stream);

/* Batch form (the `batch` argument of SURVEY.md section 8b's proposed ABI; the reference loops over patients in Python,
 * examples/time_integrated_dose.py): `batch` independent volume sets of the plan's shape through the same plan, tables and
 * cached spectrum in ONE call.  h_act holds batch * T device pointers ([b][t] order), h_dose `batch` output pointers,
 * h_density NULL or `batch` pointers (NULL entries = no correction for that volume).  The volumes run back to back on
 * `stream` (they share the plan's work buffer); the launches chain through programmatic dependent launch, so the first
 * pass of volume b + 1 is resident while the last pass of volume b drains.  Measured (C4, 256^3, T = 4): the per-patient
 * time equals the single-volume call's - the kernels, not the launches, bound it. */
int pvd_conv_execute_batch(pvd_plan* plan, const float* const* h_act, const float* h_weights, int T,
                           const float* const* h_density, float rho_ref, float rho_min, float rho_cut, float scale,
                           float* const* h_dose, int batch, void* stream);

/* Split form of pvd_conv_execute for the z-slab decomposition (no reference counterpart - the
 * reference is single-process).  The plane-local forward passes (z real-to-complex with the time-weighted sum, y forward)
 * of input planes [plane_lo, plane_hi) are independent of every other plane, so a rank can run them on its own planes
 * while the halo planes are still arriving over NVLink, then on the halo planes, and finally call pvd_conv_finish for the
 * x pass, the inverse passes and the density epilogue.  h_act[t] points at plane 0 of volume t (as in pvd_conv_execute).
 *   dose = gain * conv(sum_t w_t act_t, kernel) / max(rho, rho_min)   (gain = scale * rho_ref; no density: gain = scale)
 * pvd_conv_execute(...) == pvd_conv_forward_planes(.., gain, 0, n0, ..) followed by pvd_conv_finish(..).  FFT algorithm only;
 * every input plane must have been forwarded once before pvd_conv_finish. */
int pvd_conv_forward_planes(pvd_plan* plan, const float* const* h_act, const float* h_weights, int T, float gain,
                            int plane_lo, int plane_hi, void* stream);
int pvd_conv_finish(pvd_plan* plan, const float* density, float rho_min, float rho_cut, float* dose, void* stream);
/* pvd_conv_finish in two steps: the passes between the plane-local ones (x forward * spectrum * x inverse, y inverse), then
 * the plane-local output pass (z complex-to-real, crop, density epilogue) of OUTPUT planes [plane_lo, plane_hi) - so that an
 * end-to-end caller can feed the density map plane block by plane block while earlier blocks of the dose map already
 * travel back to the host (the link is full duplex).  density / dose point at plane 0 of the full volumes. */
int pvd_conv_middle(pvd_plan* plan, void* stream);
/* The persistent kernels launch exactly as many CTAs as stay resident on the whole GPU, and a kernel launched with
 * programmatic dependent launch is resident before its predecessor has finished - a concurrent kernel of ANOTHER stream
 * (the NCCL send/recv kernel of the halo exchange) then finds no SM with room and runs only after them (measured:
 * profiles/r02_slab_overlap_probe.jsonl).  While n_sms > 0 the persistent grids of this plan leave n_sms SMs' worth of
 * CTA slots free; 0 restores the full grids. */
int pvd_plan_reserve_sms(pvd_plan* plan, int n_sms);
int pvd_conv_output_planes(pvd_plan* plan, const float* density, float rho_min, float rho_cut, float* dose, int plane_lo,
                           int plane_hi, void* stream);

/* Stream-ordered 32-bit flags for the peer-memory halo exchange (no reference counterpart): once the work already
 * enqueued on `stream` is done, write `value` to *d_flag (which may live in ANOTHER GPU's memory, mapped through CUDA IPC);
 * hold `stream` until *d_flag >= value.  Driver stream-memory operations: no kernel, no SM - the persistent FFT kernels
 * leave no room for one (see pvd_plan_reserve_sms).  d_flag must be 4-byte aligned memory of the CURRENT device (the
 * driver rejects an IPC-mapped peer address, measured); a flag in a peer's memory is raised by copying a local word there
 * with pvd_copy_async. */
int pvd_stream_write_flag(void* d_flag, uint32_t value, void* stream);
int pvd_stream_wait_flag_geq(void* d_flag, uint32_t value, void* stream);
/* Stream-ordered copy between any two device addresses of the unified address space, peer (IPC-mapped) memory included:
 * cudaMemcpyAsync, i.e. a copy engine over NVLink, no SM.  The halo planes of the slab decomposition and the 4-byte flags
 * of their handshake travel this way (no reference counterpart). */
int pvd_copy_async(void* dst, const void* src, size_t bytes, void* stream);

int pvd_plan_destroy(pvd_plan* plan);

/* Measurement hook (bench.py / profiles): when enabled, every pvd_conv_execute brackets each of its
 * kernel launches with CUDA events on the execute stream.  pvd_plan_get_pass_times synchronises the
 * last recorded events and returns the per-launch durations (ms) and the bytes each launch moves
 * through HBM by construction; names[i] points to a static string.  Returns the number of launches. */
#define PVD_MAX_PASSES 8
int pvd_plan_set_profiling(pvd_plan* plan, int enable);
int pvd_plan_get_pass_times(pvd_plan* plan, float* ms, double* hbm_bytes, const char** names, int cap);

/* Device-side watchdogs: the kernels that wait on a TMA bulk copy (direct convolution, y passes) give up after
 * ~2 s and raise a flag in the workspace instead of hanging the GPU on a bad descriptor.  This call synchronises
 * `stream`, reads the flags and returns PVD_OK or PVD_ERR_CUDA (text in pvd_last_error()); the results of an
 * execute whose flag is raised are invalid.  The reference has no counterpart (NumPy raises synchronously). */
int pvd_plan_check_device_errors(pvd_plan* plan, void* stream);

/* ---- A5/A6/A10: dose voxel kernel evaluated on the image grid ----
 * Device evaluation of the radial dose-point-kernel form shared by the reference's generators
 * (Y90KernelGenerator.generate_kernel data/dose_kernels/y90_kernel.py:20-55,93-140;
 *  Lu177KernelGenerator.generate_kernel data/dose_kernels/lu177_kernel.py:52-86,129-184;
 *  Ga68KernelGenerator data/dose_kernels/ga68_kernel.py:20-83):
 *    k(r) = scaling * [ sum_b beta_amp[b] (1 - r/beta_range[b])^2 exp(-2 r/beta_range[b]) [r <= beta_range[b]]
 *                       + sum_p phot_amp[p] exp(-phot_mu[p] r / 10) / (4 pi r^2) [r > 0] ]
 * on a [g0][g1][g2] grid centred at g//2 with per-axis spacing r = ||(idx - g//2) * spacing_mm||
 * (the reference takes one isotropic voxel size; per-axis spacing is the A10 extension).  Arithmetic is
 * float64 on the device, output float32.  The photon term is 0 at r = 0 (the reference yields NaN there
 * for Y90/Ga68; Lu177 masks it the same way, lu177_kernel.py:176-182).  The nuclide/tissue constants
 * live in the host-side generators, as they do in the reference. */
#define PVD_RADIAL_MAX_TERMS 4
typedef struct pvd_radial_model {
    int n_beta, n_photon;
    double beta_range[PVD_RADIAL_MAX_TERMS]; /* mm */
    double beta_amp[PVD_RADIAL_MAX_TERMS];
    double phot_mu[PVD_RADIAL_MAX_TERMS];    /* 1/cm */
    double phot_amp[PVD_RADIAL_MAX_TERMS];
    double scaling;
} pvd_radial_model;
int pvd_kernel_eval_radial(const pvd_radial_model* model, const double spacing_mm[3], const int g[3], float* out,
                           void* stream);

/* ---- A9 helper: piecewise-linear HU -> mass density (clamped); h_knots = nk (hu, rho) pairs, nk <= 32 ---- */
int pvd_hu_to_density_f32(const float* hu, const float* h_knots, int nk, float* rho, size_t n, void* stream);
int pvd_hu_to_density_i16(const int16_t* hu, const float* h_knots, int nk, float* rho, size_t n, void* stream);

/* ---- Host-buffer staging: the reference's calling convention is pageable host NumPy arrays in, a host array out
 * (core/kernel_convolution.py:48-76; the examples pass float64 np.zeros volumes).  A stager owns a ring of pinned chunks
 * and a pool of host threads (csrc/host_stage.cuh).  `h_` pointers are HOST memory (pageable or pinned).
 *  - pvd_stage_h2d: n elements of `dtype` from h_src to d_dst on `stream`.  PVD_DTYPE_F64 is narrowed to float32 by the
 *    host threads (half the bytes over the link); F32 / I16 / U16 arrive as they are.  Returns when h_src has been read
 *    completely (the caller may reuse it) and every chunk copy is enqueued on `stream`.
 *  - pvd_stage_d2h: n float32 from d_src (after the work already enqueued on `stream`) to h_dst as float32 or float64
 *    (the reference returns float64); returns when h_dst is complete.
 * threads / chunk_bytes / ring_chunks = 0 pick the defaults (min(8, usable CPUs), 4 MiB, 2 * threads + 2). ---- */
#define PVD_DTYPE_F32 0
#define PVD_DTYPE_F64 1
#define PVD_DTYPE_I16 2
#define PVD_DTYPE_U16 3
typedef struct pvd_stager pvd_stager;
int pvd_stager_create(pvd_stager** out, int threads, size_t chunk_bytes, int ring_chunks);
int pvd_stager_destroy(pvd_stager* s);
int pvd_stage_h2d(pvd_stager* s, const void* h_src, int dtype, void* d_dst, size_t n, void* stream);
int pvd_stage_d2h(pvd_stager* s, const float* d_src, void* h_dst, int dtype, size_t n, void* stream);

/* 16-bit stored activity (the PET DICOM pixel data the reference reads and rescales on the host, io/dicom.py:27-47)
 * -> float32 activity on the device: out = slope * stored + intercept.  The volume crosses the link at 2 bytes/voxel. */
int pvd_i16_to_f32(const void* d_in, int is_unsigned, float slope, float intercept, float* d_out, size_t n, void* stream);

/* ---- A3: ActivitySampler._trapezoid_integration (core/activity_sampler.py:69-79) and the missing
 * integrate_dose_rates (core/dose_calculator.py:138): out = sum_t h_weights[t] * vol[t]. T <= 16. ---- */
int pvd_weighted_sum(const float* const* h_vol, const float* h_weights, int T, float* out, size_t n, void* stream);

/* ---- interpolate_timepoints (core/utils.py:154-191) and every other fixed linear combination of the sampled volumes:
 * out[j] = sum_t h_W[j * T + t] * vol[t], j < J <= 16, T <= 16, in ONE pass (each volume read once, each output written
 * once).  scipy's interp1d of kind 'linear' / 'cubic' / 'previous' - what the reference calls - is linear in the sampled
 * volumes with weights that depend on the time points only; the host computes them
 * (pyvoxeldosimetry_b200.core.utils.interpolation_weights).  A zero weight skips the product (the volume is not even
 * loaded when no output uses it), a NaN weight yields NaN ('previous' before the first sample).  h_W NULL = all ones.
 * An output may BE one of the inputs (same pointer: every thread reads its voxels of all volumes before it writes them);
 * partial overlap is not allowed.  pvd_weighted_sum is the J = 1 case. ---- */
int pvd_weighted_combine(const float* const* h_vol, int T, const float* h_W, float* const* h_out, int J, size_t n, void* stream);

/* ---- A11: TimeCurveFitting._calculate_accumulated_dose (time_integration/curve_fitting.py:74-84):
 * out = A0 / lambda * (1 - exp(-lambda * t_limit)) elementwise. ---- */
int pvd_monoexp_integral(const float* A0, const float* lambda, float t_limit, float* out, size_t n, void* stream);

/* Standalone density scaling (same formula as the fused epilogue), in place allowed. */
int pvd_density_scale(const float* dose, const float* density, float rho_ref, float rho_min, float rho_cut,
                      float scale, float* out, size_t n, void* stream);

/* ======== steps either side of the convolution ======== */

/* ---- TimeCurveFitting.fit_time_activity_curve (time_integration/curve_fitting.py:19-65): per-voxel weighted
 * least-squares fit of A0 * exp(-lambda t) to T activity volumes (the reference calls scipy.optimize.curve_fit
 * once per voxel with p0 = [y(t_0), lambda0], sigma = 1/weight_factors), fused with the closed-form integral
 * accumulated = A0/lambda * (1 - exp(-lambda * t_limit)) (:74-84).  h_vol = HOST array of T device pointers,
 * h_times / h_weights = HOST arrays of T floats (weights NULL = 1).  A0 / lambda / accumulated are device
 * outputs of n floats each; any of them may be NULL.  2 <= T <= 16.  Voxels whose fit is not finite get
 * [0, lambda0] like the reference's except-branch (:58-59). */
int pvd_monoexp_fit(const float* const* h_vol, const float* h_times, const float* h_weights, int T, float lambda0,
                    float t_limit, float* A0, float* lambda, float* accumulated, size_t n, void* stream);

/* ---- TissueComposition (tissue/composition.py:48-93) + the A9 density map, one pass over a float32 HU volume:
 *  - metal_threshold finite: voxels above it are replaced by gaussian_filter(volume with those voxels zeroed,
 *    sigma = 1, radius 4, 'reflect') evaluated at that voxel (_handle_artifacts :73-93); +INFINITY disables;
 *  - corrected (nullable): the artifact-handled HU volume;
 *  - rho (nullable): piecewise-linear HU -> density, h_knots = nk (hu, rho) pairs as pvd_hu_to_density_f32;
 *  - labels (nullable): one byte per voxel, bit c set when h_ranges[2c] <= HU <= h_ranges[2c+1]
 *    (calculate_composition :63-67), nr <= 8 ranges. */
int pvd_ct_prepare(const float* hu, const int n[3], float metal_threshold, const float* h_knots, int nk,
                   const float* h_ranges, int nr, float* corrected, float* rho, unsigned char* labels, void* stream);

/* ---- calculate_dvh (core/utils.py:233-262), on the device so the dose map need not leave HBM.
 * mask: device array of n uint8 (mask_is_f32 = 0) or float32 (1); a voxel is in the ROI when mask > 0.
 * pvd_roi_minmax: min / max / count of the ROI doses -> host (synchronises `stream`); d_scratch16 = 16 bytes of
 * device memory.  pvd_dvh_histogram: counts per bin with numpy.histogram's uniform-bin rule evaluated in float32
 * against the device edge array d_edges[bins + 1] (the host builds it with numpy.histogram_bin_edges, as
 * np.histogram does); d_hist[bins] (uint64, device) is zeroed by the call. */
int pvd_roi_minmax(const float* dose, const void* mask, int mask_is_f32, size_t n, void* d_scratch16, float* h_min,
                   float* h_max, unsigned long long* h_count, void* stream);
int pvd_dvh_histogram(const float* dose, const void* mask, int mask_is_f32, size_t n, const float* d_edges, int bins,
                      float first_edge, float last_edge, unsigned long long* d_hist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVDOSE_H */
