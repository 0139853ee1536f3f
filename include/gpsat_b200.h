/* gpsat_b200 -- C ABI of the B200-native batched local-expert GPR engine.
 *
 * Drop-in boundary for the hot path of CPOMUCL/GPSat (citations relative to the reference tree):
 * the per-expert loop body of LocalExpertOI.run (GPSat/local_experts.py:930-1260) as reached
 * through the BaseGPRModel interface (GPSat/models/base_model.py:17-448) and its GPflow
 * implementation (GPSat/models/gpflow_models.py:26-663).  The reference is pure Python and has
 * no FFI of its own; the binding a maintainer adds is a ctypes stub.
 *
 * Conventions
 *   - plain pointers and sizes only; all *_dev pointers are CUDA device pointers, everything
 *     else is host memory.  All arithmetic is IEEE float64.
 *   - every entry point returns 0 on success, a negative GPSAT_E* code or a positive cudaError_t.
 *     No C++ exception crosses the boundary.  gpsat_last_error() gives a message.
 *   - experts are passed CSR style: offsets[E+1] into row-major coords[sumN][D] / obs[sumN].
 *   - hyper-parameter vectors have stride GPSAT_MAXP: lengthscales[D], kernel_variance,
 *     likelihood_variance (constrained values, like get_parameters()).
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls that return results
 *     to host memory synchronise that stream before returning.
 */
#ifndef GPSAT_B200_H
#define GPSAT_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSAT_MAXD 4
#define GPSAT_MAXP 6
#define GPSAT_SEL_MAXTERMS 8

#define GPSAT_EINVAL (-1)   /* bad argument */
#define GPSAT_ENOMEM (-2)   /* workspace does not fit the memory budget */
#define GPSAT_ENOGPU (-3)   /* no CUDA device / wrong architecture */
#define GPSAT_ELIMIT (-4)   /* the optimiser's round limit was reached with experts still running (status 0) */
#define GPSAT_ESYNC (-5)    /* a Cholesky panel CTA timed out waiting for its slot's diagonal block (lost
                               inter-CTA flag): the affected evaluations were treated as f = +inf; results of the
                               call are complete but must not be trusted silently */

/* kernel ids: gpflow.kernels.{Matern32, Matern52, Matern12|Exponential, SquaredExponential|RBF}
 * (gpflow_models.py:116-135) */
enum { GPSAT_MATERN32 = 0, GPSAT_MATERN52 = 1, GPSAT_MATERN12 = 2, GPSAT_RBF = 3 };

/* L-BFGS termination, mirrors scipy L-BFGS-B messages (gpflow_models.py:317-329):
 * 1, 2 -> optimise_parameters() returns True; 3, 4, 5 -> False */
enum { GPSAT_OPT_RUNNING = 0, GPSAT_OPT_CONV_PGTOL = 1, GPSAT_OPT_CONV_FTOL = 2,
       GPSAT_OPT_MAXITER = 3, GPSAT_OPT_MAXFUN = 4, GPSAT_OPT_ABNORMAL = 5 };

typedef struct gpsat_handle gpsat_handle;

/* One batch of experts resident in device memory.  Replaces the (df_local, coords_col, obs_col,
 * coords_scale, obs_scale, obs_mean) arguments of BaseGPRModel.__init__ (base_model.py:83-245). */
typedef struct {
  int n_experts;
  int D;                          /* coordinate dimension, <= GPSAT_MAXD */
  int kernel_id;
  int obs_mean_local;             /* 1 <=> obs_mean="local" (base_model.py:195-198) */
  const long long* offsets_host;  /* [E+1] */
  const long long* offsets_dev;   /* [E+1] */
  const double* coords_dev;       /* [sumN][D] raw coordinates */
  const double* obs_dev;          /* [sumN] */
  double coords_scale[GPSAT_MAXD];
  double obs_scale;
  double* obs_mean_out_dev;       /* [E] or NULL: the mean that was subtracted (-> "f_bar") */
} gpsat_batch;

/* Parameter transforms: gpflow defaults (softplus; likelihood variance softplus + 1e-6) or the
 * tfp Sigmoid(low, high) installed by set_*_constraints (gpflow_models.py:416-494,592-628). */
typedef struct {
  int kind[GPSAT_MAXP];      /* 0: softplus + low, 1: sigmoid(low, high) */
  double low[GPSAT_MAXP];
  double high[GPSAT_MAXP];
  int trainable[GPSAT_MAXP]; /* 0 <=> listed in fixed_params (gpflow_models.py:275-288) */
} gpsat_transforms;

/* scipy L-BFGS-B options as used by gpflow.optimizers.Scipy (defaults in gpsat_default_opts) */
typedef struct {
  int maxcor, maxiter, maxfun, maxls;
  double ftol, gtol;
} gpsat_opt_options;

/* local_select / max_dist predicate list (dataloader.py:2405-2444; prediction_locations.py:18-43)
 * type 0: table[col[0]] <comp> ref[rcol[0]] + val ; comp: 0 >=, 1 >, 2 ==, 3 <, 4 <=
 * type 1: sum_j (table[col[j]] - ref[rcol[j]])^2 <= val*val      (KDTree.query_ball_point)
 * type 2: strict per-dimension and squared-L2 test against val   (_max_dist_bool) */
typedef struct {
  int type, ncol, comp, pad_;
  int col[4];
  int rcol[4];
  double val;
} gpsat_sel_term;
typedef struct {
  int nterms, pad_;
  gpsat_sel_term t[GPSAT_SEL_MAXTERMS];
} gpsat_sel_spec;

const char* gpsat_last_error(void);
int gpsat_version(void);

/* lifetime: owns cached device workspaces; mem_budget_bytes = 0 -> 70% of free device memory */
int gpsat_create(gpsat_handle** out, int device, size_t mem_budget_bytes);
int gpsat_destroy(gpsat_handle* h);
void gpsat_default_opts(gpsat_opt_options* o);

/* S2 / S3: two calls.  counts_dev[E] <- matches; then (after an exclusive scan by the caller)
 * idx_dev[offsets_dev[e] ...] <- matching row indices in ascending order.
 * table_dev: [ncols][n] column-major; refs_dev: [E][nrefcols] row-major. */
int gpsat_select_count(const gpsat_sel_spec* spec, const double* table_dev, long long n,
                       const double* refs_dev, int nrefcols, int n_experts,
                       long long* counts_dev, void* stream);
int gpsat_select_fill(const gpsat_sel_spec* spec, const double* table_dev, long long n,
                      const double* refs_dev, int nrefcols, int n_experts,
                      const long long* offsets_dev, int* idx_dev, void* stream);

/* Grid-bucketed S2 / S3 (same predicates, same bit-exact index sets, ascending row order): rows are
 * binned once on the two columns of the spec's ball (type 1) or max_dist (type 2) term into square
 * cells of edge `cell` > radius; an expert then only tests its 3 x 3 cell neighbourhood.
 *   gpsat_bucket_build pass 0: counts_dev[ncx*ncy] += rows per cell (caller zeroes it, then exclusive-scans
 *                              it into start_dev[ncx*ncy + 1] and zeroes counts_dev again)
 *                      pass 1: order_dev[n] <- row ids grouped by cell (counts_dev is the cursor)
 *   gpsat_select_bucket: counts_dev != NULL -> matches per expert; counts_dev == NULL -> fill idx_dev at
 *                        offsets_dev (max_count = largest per-expert count, <= 32768). */
typedef struct {
  double x0, y0, cell;
  int ncx, ncy;
} gpsat_cell_grid;
int gpsat_bucket_build(const gpsat_sel_spec* spec, const gpsat_cell_grid* grid, const double* table_dev,
                       long long n, int pass, int* counts_dev, const long long* start_dev, int* order_dev,
                       void* stream);
int gpsat_select_bucket(const gpsat_sel_spec* spec, const gpsat_cell_grid* grid, const double* table_dev,
                        long long n, const double* refs_dev, int nrefcols, int n_experts,
                        const long long* start_dev, const int* order_dev, int max_count, long long* counts_dev,
                        const long long* offsets_dev, int* idx_dev, void* stream);

/* pack the selected rows into the CSR batch layout: coords_dev[total][D] <- table columns
 * coord_cols[0..D-1]; obs_dev[total] <- column obs_col (obs_col < 0: skipped).  Replaces
 * df.loc[select, :] + data[coords_col].values / data[obs_col].values
 * (dataloader.py:2447; base_model.py:148-149). */
int gpsat_gather_rows(const double* table_dev, long long n, const int* idx_dev, long long total, int D,
                      const int* coord_cols, int obs_col, double* coords_dev, double* obs_dev, void* stream);

/* prediction coordinates out_dev[total][D] of every expert: column table_cols[d] of the selected
 * prediction-location rows, or (table_cols[d] < 0) the expert's own refs[e][ref_cols[d]]
 * (PredictionLocations._from_dataframe, prediction_locations.py:258-271). */
int gpsat_gather_pred(const double* table_dev, long long n, const double* refs_dev, int nrefcols,
                      int n_experts, const long long* offsets_dev, const int* idx_dev, int D,
                      const int* table_cols, const int* ref_cols, double* out_dev, void* stream);

/* K1: dense kernel matrix K(X1, X2) [n1][n2] row-major (+ likelihood variance on the diagonal
 * when add_noise); coordinates are raw, divided by coords_scale inside. theta_dev[GPSAT_MAXP]. */
int gpsat_kernel_matrix(const double* x1_dev, int n1, const double* x2_dev, int n2, int D,
                        int kernel_id, const double* theta_dev, int add_noise, double* k_dev,
                        void* stream);

/* L1 + G1: -LML (get_objective_function_value, gpflow_models.py:334-337) and, when
 * grad_dev != NULL, d(-LML)/d(theta) for every expert at theta_dev[E][GPSAT_MAXP].
 * fail_dev[E] (optional) is 1 where the Cholesky met a non-positive pivot (f = +inf). */
int gpsat_gpr_eval(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                   double* f_dev, double* grad_dev, void* stream);

/* P1: optimise_parameters for every expert (gpflow_models.py:290-329).
 * theta0_dev: start values after set_parameter_constraints' move_within_tol (host logic).
 * Outputs per expert: optimised theta, final -LML, termination status, iterations, evaluations. */
int gpsat_gpr_optimise(gpsat_handle* h, const gpsat_batch* b, const double* theta0_dev,
                       const gpsat_transforms* tr, const gpsat_opt_options* opts,
                       double* theta_out_dev, double* fobj_out_dev, int* status_out_dev,
                       int* nit_out_dev, int* nfev_out_dev, void* stream);

/* F1: predict (gpflow_models.py:186-273) at pred_coords[sumP][D] (raw, CSR by pred_offsets):
 * f* (fmean), f*_var (fvar), y_var = f*_var + likelihood variance; fobj_dev (optional) gets -LML
 * at theta (the value run() stores as objective_value, local_experts.py:1135). */
int gpsat_gpr_predict(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                      const long long* pred_offsets_host, const long long* pred_offsets_dev,
                      const double* pred_coords_dev, double* fmean_dev, double* fvar_dev,
                      double* yvar_dev, double* fobj_dev, void* stream);

/* F1 with full_cov=True (gpflow_models.py:245-263) for the FIRST expert of the batch:
 * fmean_dev[P], fcov_dev[P][P] (row-major posterior covariance of f*). */
int gpsat_gpr_predict_cov(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                          const double* pred_coords_dev, int P, double* fmean_dev, double* fcov_dev,
                          void* stream);

/* SG1: sparse GPR (GPflowSGPRModel, gpflow_models.py:666-901; gpflow.models.SGPR with Kuu jitter 1e-6).
 * `data` is the usual CSR batch (N observations per expert); the inducing points are a second CSR
 * (M per expert, raw coordinates, fixed: train_inducing_points=False is the reference's default). */
typedef struct {
  gpsat_batch data;
  const long long* z_offsets_host;   /* [E+1] */
  const long long* z_offsets_dev;    /* [E+1] */
  const double* z_coords_dev;        /* [sumM][D] */
} gpsat_sgpr_batch;

/* f_dev[E] = -ELBO (the training loss gpflow minimises; get_objective_function_value() returns +ELBO,
 * gpflow_models.py:860-862) and grad_dev[E][GPSAT_MAXP] = d(-ELBO)/d(theta) (optional). */
int gpsat_sgpr_eval(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta_dev, double* f_dev,
                    double* grad_dev, void* stream);
/* optimise_parameters(train_inducing_points=False) (gpflow_models.py:864-901) */
int gpsat_sgpr_optimise(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta0_dev,
                        const gpsat_transforms* tr, const gpsat_opt_options* opts, double* theta_out_dev,
                        double* fobj_out_dev, int* status_out_dev, int* nit_out_dev, int* nfev_out_dev,
                        void* stream);
/* predict (gpflow SGPR.predict_f / predict_y); fobj_dev (optional) gets -ELBO at theta */
int gpsat_sgpr_predict(gpsat_handle* h, const gpsat_sgpr_batch* sb, const double* theta_dev,
                       const long long* pred_offsets_host, const long long* pred_offsets_dev,
                       const double* pred_coords_dev, double* fmean_dev, double* fvar_dev, double* yvar_dev,
                       double* fobj_dev, void* stream);

/* test hook: dense lower factor L and its inverse X (both (nb*64)^2 row-major, nb = n/64+1, of
 * the augmented matrix) of expert 0 of the batch at theta_dev. */
int gpsat_debug_factor(gpsat_handle* h, const gpsat_batch* b, const double* theta_dev,
                       double* l_dense_dev, double* x_dense_dev, void* stream);

/* counters: kernels launched by this library since the handle was created (bench: gpu_launches),
 * and CUDA-event time spent in the phases of an objective evaluation when profiling is enabled:
 * potrf = k_potrf_* + k_quad (the batched Cholesky), trtri, lauum = k_lauum2, other = k_finalize2,
 * build = k_build (kernel matrix), trace = k_grad_trace; flops_* = sum of N^3/3 over the slots evaluated */
long long gpsat_launch_count(const gpsat_handle* h);
/* Slot plan of the last batched call on this handle: resident experts (slots), 64-row blocks of the largest
 * matrix, bytes of workspace per slot and the memory budget the plan was fitted to.  Any pointer may be NULL. */
int gpsat_last_plan(const gpsat_handle* h, int* slots, int* nbmax, size_t* bytes_per_slot, size_t* budget_bytes);
/* Flag-wait timeouts of k_potrf_panel since the handle was created (0 on a healthy device; see GPSAT_ESYNC). */
long long gpsat_sync_timeouts(gpsat_handle* h);
int gpsat_set_profiling(gpsat_handle* h, int enabled);
int gpsat_get_profile(gpsat_handle* h, double* ms_potrf, double* ms_trtri, double* ms_lauum,
                      double* ms_other, double* flops_potrf, double* flops_trtri, double* flops_lauum,
                      double* ms_build, double* ms_trace);

/* Post-processing either side of the hot path.
 * gpsat_gaussian_smooth replaces gaussian_2d_weight (GPSat/postprocessing.py:22-52) as called by
 *   smooth_hyperparameters (postprocessing.py:269-292): out[i] = sum_j w_ij v_j / sum_j w_ij over the rows j of the
 *   segment of query i, w_ij = exp(-(((x_j - qx_i)/l_x)^2 + ((y_j - qy_i)/l_y)^2) / 2); NaN values are skipped, a
 *   segment with no finite value yields NaN; vmin / vmax (host pointers, NULL = none) clip the values first.
 *   Rows are grouped by segment: seg_off[G+1] (int64), seg_of_query[n_query] (int32).
 * gpsat_weighted_groups replaces get_weighted_values (GPSat/utils.py:2081-2214) and the Gaussian glue
 *   (postprocessing.py:447-577): ref / to are [nd][n] column-major, vals [ncol][n]; group g = source rows
 *   order[group_off[g] .. group_off[g+1]); w = exp(-(|ref - to|^2 / lengthscale^2) / 2);
 *   out [ncol + 1][G] = weighted mean per column, last row = sum of w. */
int gpsat_gaussian_smooth(const double* qx_dev, const double* qy_dev, const int* seg_of_query_dev,
                          long long n_query, const double* x_dev, const double* y_dev, const double* vals_dev,
                          const long long* seg_off_dev, double l_x, double l_y, const double* vmin,
                          const double* vmax, double* out_dev, void* stream);
int gpsat_weighted_groups(const double* ref_dev, const double* to_dev, int nd, const double* vals_dev, long long n,
                          int ncol, const long long* order_dev, const long long* group_off_dev,
                          long long n_groups, double lengthscale, double* out_dev, void* stream);

/* Upstream binning: replaces scipy.stats.binned_statistic_2d as called by DataPrep.bin_data
 * (GPSat/dataprepper.py:230-407) for every by_cols group of DataPrep.bin_data_by (dataprepper.py:23-228) in one launch.
 * Bin numbers follow scipy: searchsorted(edges, v, side="right"), a value that rounds onto the last edge
 * (around(v, decimal), decimal = int(-log10(min edge step)) + 6: round_scale = 10^|decimal|, round_div = decimal < 0)
 * belongs to the last bin; out-of-range rows are dropped.  y_dev = NULL: 1-D binning.  sum_dev (fp64) and count_dev
 * (uint64), both [n_groups][n_x_edges - 1][n_y_edges - 1] and zeroed by the caller, receive the per-bin sum of vals
 * and the number of rows; group_dev (int32, NULL = one group) is the row's group. */
int gpsat_bin_accumulate(const double* x_dev, const double* y_dev, const double* vals_dev, const int* group_dev,
                         long long n, const double* x_edges_dev, int n_x_edges, double x_round_scale, int x_round_div,
                         const double* y_edges_dev, int n_y_edges, double y_round_scale, int y_round_div,
                         int n_groups, double* sum_dev, unsigned long long* count_dev, void* stream);

/* Second pass over the same rows for scipy's statistic = "std" | "min" | "max" (GPSat/dataprepper.py:359 as called
 * with bin_statistic=["mean", "std", "count"] in examples/bin_data.py:165): with sum_dev / count_dev of
 * gpsat_bin_accumulate, ssd_dev (zeroed by the caller) receives the per-bin sum of (v - mean)^2 -- np.std's two-pass
 * form, std = sqrt(ssd / count) -- and min_dev / max_dev (initialised to +inf / -inf) the per-bin extrema.  Any of the
 * three outputs may be NULL; sum_dev / count_dev are needed only with ssd_dev.  Same row -> bin map as the first pass. */
int gpsat_bin_spread(const double* x_dev, const double* y_dev, const double* vals_dev, const int* group_dev,
                     long long n, const double* x_edges_dev, int n_x_edges, double x_round_scale, int x_round_div,
                     const double* y_edges_dev, int n_y_edges, double y_round_scale, int y_round_div, int n_groups,
                     const double* sum_dev, const unsigned long long* count_dev, double* ssd_dev, double* min_dev,
                     double* max_dev, void* stream);

/* roofline denominator for the factorisation kernels: FP64 mma.sync (DMMA m8n8k4) issue rate of
 * this GPU measured with register-resident accumulator chains (no memory traffic). */
int gpsat_dmma_peak(int device, int iters, double* tflops_out, double* ms_out);

/* isolated core measurements (profiles/): which 0..3 = register DMMA chains with 1/2/4/8
 * accumulators per warp and `param` CTAs of 8 warps per SM (nk = iterations); 10 = 64x64 core,
 * 11/12/13 = 128x128 core (NT / TN / NN operand orientations) with param = 0 shared-memory resident,
 * 1 streaming private tiles (HBM), 2 all CTAs on the same tiles (L2); nk = 64-deep k steps. */
int gpsat_microbench(int device, int which, int param, int nk, double* tflops_out);

/* host-side (CPU) entry to the SAME L-BFGS state machine the device runs, for CPU unit tests of
 * the optimiser logic against scipy (no GPU needed). */
size_t gpsat_lbfgs_state_bytes(void);
void gpsat_lbfgs_init_host(void* state, const double* x0, int n);
int gpsat_lbfgs_tell_host(void* state, const gpsat_opt_options* o, double f, const double* g,
                          double* x_next, int* nit, int* nfev);

#ifdef __cplusplus
}
#endif
#endif /* GPSAT_B200_H */
