/*
 * openfly_b200 -- C ABI of the B200-native force-evaluation path for openFLY.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types. The C++20 shim in
 * include/openfly_b200/fly.hpp re-creates the reference's class surface (fly::potential::Generic,
 * fly::neigh::List, fly::minimise::LBFGS, fly::saddle::Dimer) on top of these entry points, and
 * openfly_b200/capi.py binds the same symbols through ctypes for the tests and bench.py.
 *
 * Reference interfaces replaced (paths relative to the openFLY tree):
 *   ofly_tables_*          potential::DataEAM            src/potential/EAM/data.cpp:67-153,
 *                                                        include/libfly/potential/EAM/data.hpp:59-116
 *   ofly_neigh_*           neigh::List                   include/libfly/neigh/list.hpp:86-161, src/neigh/list.cpp:37-165
 *   ofly_eam_*             potential::Generic / EAM      include/libfly/potential/generic.hpp:79-136, src/potential/EAM/eam.cpp:33-233
 *   ofly_lbfgs_minimise    minimise::LBFGS::minimise     include/libfly/minimise/LBFGS/lbfgs.hpp:96-118, src/minimise/LBFGS/lbfgs.cpp:29-115
 *   ofly_dimer_find_sp     saddle::Dimer::find_sp        include/libfly/saddle/dimer.hpp:107-155, src/saddle/dimer.cpp:28-171
 *                          (+ saddle::Rotor::eff_gradient src/saddle/rotor.cpp:49-164)
 *
 * Data layout at the boundary is the reference's: every vector property is one contiguous
 * double[3N] with xyz interleaved per atom (include/libfly/system/adaptor.hpp:70-78); TypeID is
 * uint32[N]; Frozen is one byte per atom. A "batch" is R independent replicas of one N-atom
 * system (same box, types and frozen flags -- what saddle::Master::find_batch launches,
 * src/saddle/find.cpp:684-755); batched arrays are [R][3N], replica-major.
 *
 * Every call returns an ofly_status. Errors never abort: device-side faults (table overrun,
 * neighbour-capacity overflow) are recorded per replica and reported by the call that saw them.
 * Pointers named *_host are host memory, *_dev are device memory on the context's device.
 */
#ifndef OPENFLY_B200_H
#define OPENFLY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ofly_status {
  OFLY_OK = 0,
  OFLY_ERR_INVALID = 1,       /* contract violation: fly::RuntimeError in the reference (core.hpp:89-138) */
  OFLY_ERR_OUT_OF_BOUNDS = 2, /* Spline::OutOfBounds (include/libfly/utility/spline.hpp:46-51): recoverable */
  OFLY_ERR_CUDA = 3,          /* CUDA runtime failure */
  OFLY_ERR_CAPACITY = 4       /* an internal buffer bound was exceeded even after regrowing */
} ofly_status;

typedef struct ofly_ctx ofly_ctx;       /* one device + one stream; calls on a context are serialised */
typedef struct ofly_tables ofly_tables; /* immutable, ref-counted EAM tables resident on the device */
typedef struct ofly_batch ofly_batch;   /* R replicas x N atoms: list-space positions, neighbour lists, workspaces */

/* Message of the last error raised on this host thread. */
const char *ofly_last_error(void);
/* Library version string. */
const char *ofly_version(void);

int ofly_ctx_create(int device, ofly_ctx **out);
/* Destroy every batch and release every table handle made on a context before the context itself. */
void ofly_ctx_destroy(ofly_ctx *ctx);
/* The context's cudaStream_t (so a caller can time it with events or order work after it). */
void *ofly_ctx_stream(ofly_ctx *ctx);
int ofly_ctx_synchronize(ofly_ctx *ctx);
/* Number of kernels this context has launched so far. */
long ofly_ctx_launch_count(ofly_ctx *ctx);
/* Number of lock-step ticks (one gradient evaluation of every still-active replica each) the search calls have run. */
long ofly_ctx_tick_count(ofly_ctx *ctx);

/* ---- EAM tables: potential::DataEAM -------------------------------------------------------- */

/* Parse a LAMMPS setfl eam/fs file and build the natural cubic splines (data.cpp:67-153,
 * src/utility/spline.cpp:27-103), then upload the knots. */
int ofly_tables_load_setfl(ofly_ctx *ctx, const char *path, ofly_tables **out);
/* The same from memory: DataEAM(std::istream&) (data.cpp:67) for hosts that hold the file's text. */
int ofly_tables_load_setfl_mem(ofly_ctx *ctx, const char *text, size_t len, ofly_tables **out);
/* From already-splined knots (a host that owns a fly::potential::DataEAM hands its Spline coefficients over):
 * names/mass: [n_types]; F_abcd: [n_types][n_knots_F][4] = {a,b,c,d} (spline.hpp:121-126); phi_abcd: [src + n*dst][n_knots_r][4]
 * (data.hpp:109-112); V_abcd: [max + n*min][n_knots_r][4] (data.hpp:114-116, unused slots ignored). */
int ofly_tables_create_from_knots(ofly_ctx *ctx, int n_types, const char *const *names, const double *mass, long n_knots_F, double d_rho,
                                  long n_knots_r, double d_r, double r_cut, const double *F_abcd, const double *phi_abcd,
                                  const double *V_abcd, ofly_tables **out);
/* Name of type i as the table declares it (line 4 of the setfl file); NULL if out of range. */
const char *ofly_tables_type_name(const ofly_tables *t, int i);
/* EAM's constructor check (include/libfly/potential/EAM/eam.hpp:50-63): the caller's TypeMap must have at least the table's
 * types, with the same names at the same ids. OFLY_ERR_INVALID with the reference's message otherwise. */
int ofly_eam_check_type_map(const ofly_tables *t, int n_types, const char *const *names);
void ofly_tables_retain(ofly_tables *t);
void ofly_tables_release(ofly_tables *t);
/* info = {n_types, n_knots_F, n_knots_r}; dinfo = {d_rho, d_r, r_cut, mass[0..n_types)} */
int ofly_tables_info(const ofly_tables *t, long info[3], double *dinfo);
/* Copy one spline's knots {a,b,c,d} x n_knots to host. kind: 0 = F_i, 1 = phi(i -> j), 2 = V_ij. */
int ofly_tables_knots(const ofly_tables *t, int kind, int i, int j, double *out_host);
/* Host-only variants (no device needed): parse + spline a setfl file and return its header / one
 * spline's knots. Same arithmetic as ofly_tables_load_setfl; used to check tables bit for bit. */
int ofly_setfl_info_host(const char *path, long info[3], double *dinfo);
int ofly_setfl_knots_host(const char *path, int kind, int i, int j, double *out_host);
/* Generic::r_cut() (generic.hpp:79). */
double ofly_eam_r_cut(const ofly_tables *t);

/* ---- batch ---------------------------------------------------------------------------------- */

/* extents: orthorhombic box lengths (system::Box with skew < 1e-4, box.hpp:63-76); periodic: per axis. */
int ofly_batch_create(ofly_ctx *ctx, ofly_tables *tables, const double extents[3], const int periodic[3], long n_atoms,
                      long n_replicas, const uint32_t *type_host, const uint8_t *frozen_host, ofly_batch **out);
void ofly_batch_destroy(ofly_batch *b);
long ofly_batch_atoms(const ofly_batch *b);
long ofly_batch_replicas(const ofly_batch *b);

/* ---- neigh::List ---------------------------------------------------------------------------- */

/* List(box, r_list) + rebuild(positions) (list.cpp:53-95) for every replica. pos: [R][3N]. */
int ofly_neigh_rebuild(ofly_batch *b, double r_list, const double *pos_host);
int ofly_neigh_rebuild_dev(ofly_batch *b, double r_list, const double *pos_dev);
/* update(x): r <- r - x for real atoms and their images (list.cpp:37-51). delta: [R][3N]. */
int ofly_neigh_update(ofly_batch *b, const double *delta_host);
int ofly_neigh_update_dev(ofly_batch *b, const double *delta_dev);
/* Stored (unfiltered) neighbour lists of one replica as (real index, image shift code) pairs,
 * shift code = (sx+1) + 3(sy+1) + 9(sz+1). counts: [N]; idx/shift: flattened, capacity cap;
 * *total receives the number of pairs (may exceed cap, then only cap were written). An atom's pairs come in stored order:
 * ascending (real index, shift code) -- the reference's own order (cell walk over its ghost array) is not reproduced, the set is. */
int ofly_neigh_export(ofly_batch *b, long replica, int *counts_host, int *idx_host, int *shift_host, long cap, long *total);
/* HyperGrid::cell_idx (hypergrid.hpp:75-99) of every real atom's canonical image. out: [N]. */
int ofly_neigh_cells(ofly_batch *b, long replica, int *cell_host);
/* List-space (canonical, updated) positions of the real atoms of one replica. out: [3N]. */
int ofly_neigh_positions(ofly_batch *b, long replica, double *pos_host);
/* Mean stored neighbours per atom over the whole batch (for the roofline's n-bar). */
int ofly_neigh_mean_count(ofly_batch *b, double *mean);

/* List::for_neighbours (list.hpp:138-154) of ONE atom, evaluated on the device: the neighbours of `atom` within r_cut as
 * (real index, r, dr = r_n - r_i with the image shift applied), in stored order; only the hits travel. idx: [cap], r: [cap],
 * dr: [cap][3]; *count receives the number of hits (may exceed cap). For iterating over ALL atoms use ofly_neigh_export once. */
int ofly_neigh_for_atom(ofly_batch *b, long replica, long atom, double r_cut, int *idx_host, double *r_host, double *dr_host, long cap,
                        long *count);

/* The same for EVERY atom of a replica in one pass -- what Catalogue::rebuild_impl / rebuild_geo_from_nl (src/env/catalogue.cpp:80-96)
 * and DetectVacancies (src/utility/lattice.cpp:69-84) loop for: CSR output, atom i's hits at [offsets[i], offsets[i+1]). offsets: [N+1]
 * (always written); idx / r: [cap], dr: [cap][3] (skipped when NULL or cap < *total: call once for the size, again with buffers). */
int ofly_neigh_within_all(ofly_batch *b, long replica, double r_cut, long *offsets_host, int *idx_host, double *r_host, double *dr_host,
                          long cap, long *total);

/* ---- potential::Generic (EAM) over the batch's current list ---------------------------------- */

/* energy (eam.cpp:33-51): out [R]. */
int ofly_eam_energy(ofly_batch *b, double *energy_host);
/* gradient (eam.cpp:53-120): out [R][3N]; frozen atoms get zero. */
int ofly_eam_gradient(ofly_batch *b, double *grad_host);
int ofly_eam_gradient_dev(ofly_batch *b, double *grad_dev);
/* hessian (eam.cpp:122-233) of one replica: dense column-major (3N)^2, lower triangle meaningful,
 * zero rows/cols for frozen atoms. mass_weight != 0 applies Hessian::mass_weight (hessian.hpp:136-143). */
int ofly_eam_hessian(ofly_batch *b, long replica, int mass_weight, double *hess_host);
/* Hessian::eigenvalues() (hessian.hpp:86-93) of that matrix, ascending, solved on the device: out [3N]. The consumers are
 * the Vineyard prefactors of saddle::Master (src/saddle/find.cpp:846-871, 1107-1135). Dense symmetric solve by cuSOLVER
 * (loaded on first use); returns OFLY_ERR_CUDA if it is not installed. */
int ofly_eam_hessian_eigenvalues(ofly_batch *b, long replica, int mass_weight, double *eig_host);

/* ---- minimise::LBFGS and saddle::Dimer, whole batch resident on the device -------------------- */

typedef struct ofly_lbfgs_options { /* minimise::LBFGS::Options, lbfgs.hpp:53-88 */
  int n;
  int iter_max;
  double f2norm;
  double skin_frac;
  double proj_tol;
  double max_trust;
  double min_trust;
  double grow_trust;
  double shrink_trust;
} ofly_lbfgs_options;

typedef struct ofly_dimer_options { /* saddle::Dimer::Options, dimer.hpp:56-98 */
  int n;
  int max_steps;
  int hist_check_freq;
  int use_history;
  double f2norm;
  double skin_frac;
  double proj_tol;
  double max_trust;
  double min_trust;
  double grow_trust;
  double shrink_trust;
  double convex_max;
} ofly_dimer_options;

typedef struct ofly_rotor_options { /* saddle::Rotor::Options, rotor.hpp:47-60 */
  int n;
  int iter_max_rot;
  int relax_in_convex;
  double delta_r;
  double theta_tol;
} ofly_rotor_options;

void ofly_lbfgs_options_default(ofly_lbfgs_options *o);
void ofly_dimer_options_default(ofly_dimer_options *o);
void ofly_rotor_options_default(ofly_rotor_options *o);

/* Per-replica work counters filled by the search calls: {grad_evals, rebuilds, updates, steps}. */
#define OFLY_NUM_COUNTERS 4

/* LBFGS::minimise for every replica. pos_in/pos_out/grad_out: [R][3N].
 * status[r]: 0 converged, 1 iter_max (the reference returns this bool inverted, lbfgs.hpp:108-118),
 * -5 Spline::OutOfBounds was hit (reference: exception), -6 neighbour capacity fault.
 * counters: [R][OFLY_NUM_COUNTERS] or NULL. */
int ofly_lbfgs_minimise(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_host, double *pos_out_host,
                        double *grad_out_host, int *status_host, long *counters_host);
int ofly_lbfgs_minimise_dev(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_dev, double *pos_out_dev,
                            double *grad_out_dev, int *status_dev, long *counters_dev);

/* Dimer::find_sp for every replica. pos/axis in and out: [R][3N]; in_min: [3N]; hist_sp: [n_hist][3N].
 * exit_code[r]: Dimer::Exit (dimer.hpp:123-129: 0 success, -1 convex, -2 iter_max, -3 collision),
 * -5 Spline::OutOfBounds, -6 neighbour capacity fault. */
int ofly_dimer_find_sp(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_host,
                       const double *axis_in_host, const double *in_min_host, const double *hist_sp_host, long n_hist,
                       double theta_tol, int count_frozen, double *pos_out_host, double *axis_out_host, int *exit_code_host,
                       long *counters_host);
int ofly_dimer_find_sp_dev(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_dev,
                           const double *axis_in_dev, const double *in_min_dev, const double *hist_sp_dev, long n_hist,
                           double theta_tol, int count_frozen, double *pos_out_dev, double *axis_out_dev, int *exit_code_dev,
                           long *counters_dev);

/* Queue forms of the two calls above: n_searches start states run through the batch's R resident slots. A slot whose
 * search has finished claims the next start state on the device, inside the same graph launch, so the GPU stays full
 * until the queue is empty -- saddle::Master::find_n's `while (tot < max_searches)` loop (src/saddle/find.cpp:374-422)
 * without a barrier after every batch. All per-search arrays are [n_searches][...] and indexed by search; in_min and
 * hist_sp are shared by the whole queue (the SP cache is read-only during a call, find.cpp:732-752). What a search
 * returns does not depend on n_searches, on R, or on which searches share the device with it. n_searches == R is
 * exactly ofly_lbfgs_minimise / ofly_dimer_find_sp. */
int ofly_lbfgs_minimise_stream(ofly_batch *b, const ofly_lbfgs_options *opt, long n_searches, const double *pos_in_host,
                               double *pos_out_host, double *grad_out_host, int *status_host, long *counters_host);
int ofly_lbfgs_minimise_stream_dev(ofly_batch *b, const ofly_lbfgs_options *opt, long n_searches, const double *pos_in_dev,
                                   double *pos_out_dev, double *grad_out_dev, int *status_dev, long *counters_dev);
int ofly_dimer_find_sp_stream(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, long n_searches,
                              const double *pos_in_host, const double *axis_in_host, const double *in_min_host,
                              const double *hist_sp_host, long n_hist, double theta_tol, int count_frozen, double *pos_out_host,
                              double *axis_out_host, int *exit_code_host, long *counters_host);
int ofly_dimer_find_sp_stream_dev(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, long n_searches,
                                  const double *pos_in_dev, const double *axis_in_dev, const double *in_min_dev,
                                  const double *hist_sp_dev, long n_hist, double theta_tol, int count_frozen, double *pos_out_dev,
                                  double *axis_out_dev, int *exit_code_dev, long *counters_dev);

/* ---- saddle::Master's per-mechanism numerics, batched over the saddle points of one environment ------------------------ */

/* hessian (eam.cpp:122-233) of EVERY replica in one launch: out [R][(3N)^2], layout as ofly_eam_hessian. */
int ofly_eam_hessian_batch(ofly_batch *b, int mass_weight, double *hess_host);

typedef struct ofly_mech_options { /* the fields of saddle::Master::Options these steps read, find.hpp:112-184 */
  double nudge_frac;             /* 0.01 */
  double basin_tol;              /* 0.05 */
  double stationary_tol;         /* basin_tol / 2 */
  double hessian_eigen_zero_tol; /* 1e-6 */
  long centre;        /* geo[0][i_]: the atom the mechanism must be centred on (find.cpp:1082-1087); < 0: not checked */
  long freeze_atom;   /* m_sep_list[centre] (ofly_mech_sep_list): frozen during the two relaxations when nothing else is
                         (find.cpp:1026-1034); < 0: none */
  int num_zero_modes; /* Master::m_num_zero_modes as ofly_mech_minimum returned it */
  int skip_pathway;   /* != 0: no relaxations; E0 = energy of in_min, only the saddle's energy / Hessian rules are evaluated */
} ofly_mech_options;
void ofly_mech_options_default(ofly_mech_options *o);

typedef enum ofly_mech_status {
  OFLY_MECH_OK = 0,                    /* a mechanism: min -> sp -> min, first-order saddle */
  OFLY_MECH_MINIMISATION_FAILED = 1,   /* do_adj_min returns {} (find.cpp:1038-1050) */
  OFLY_MECH_NOT_CENTRED = 2,           /* largest displacement is not on `centre` (find.cpp:1084) */
  OFLY_MECH_UNLINKED = 3,              /* |rev - in| > basin_tol (find.cpp:779) */
  OFLY_MECH_BACK_TO_INITIAL = 4,       /* |fwd - rev| < basin_tol (find.cpp:783) */
  OFLY_MECH_MIN_TO_SP_TOO_SHORT = 5,   /* |rev - sp| < stationary_tol (find.cpp:787) */
  OFLY_MECH_SP_TO_MIN_TOO_SHORT = 6,   /* |fwd - sp| < stationary_tol (find.cpp:791) */
  OFLY_MECH_HIGHER_ORDER_SADDLE = 7,   /* a second negative mode (find.cpp:852-855) */
  OFLY_MECH_ERR_SP_BELOW_MIN = -1,     /* the reference's verify(Esp > E0) (find.cpp:807) */
  OFLY_MECH_ERR_NO_NEGATIVE_MODE = -2, /* verify(freq[0] < -tol) (find.cpp:846) */
  OFLY_MECH_ERR_ZERO_MODES = -3,       /* zero-mode count differs from the minimum's (throw, find.cpp:862-867) */
  OFLY_MECH_OUT_OF_BOUNDS = -5         /* Spline::OutOfBounds in a relaxation, an energy or the Hessian */
} ofly_mech_status;

typedef struct ofly_mech_record {
  int status;      /* ofly_mech_status */
  int count_zeros; /* zero modes of the saddle's mass-weighted Hessian */
  double i2r, r2f, r2w, w2f; /* |rev - in|, |fwd - rev|, |rev - sp|, |fwd - sp| (find.cpp:772-775) */
  double E0, Esp, Ef;        /* energies of rev, sp, fwd (find.cpp:797-805) */
  double barrier, delta;     /* Esp - E0, Ef - E0 */
  double log_prod_eigen;     /* sum of log(lambda) over modes 1.. above the tolerance: what Mechanism::kinetic_pre holds
                                before find.cpp:287; ofly_mech_prefactor turns it into the attempt frequency */
  double min_mode;           /* freq[0] */
} ofly_mech_record;

/* m_sep_list (find.cpp:223-244): for every atom the index of the atom furthest away under the minimum image. pos: [3N], out: [N]. */
int ofly_mech_sep_list(ofly_batch *b, const double *pos_host, int *sep_host);
/* Master::calc_minima_hess (find.cpp:1094-1144): mass-weighted Hessian of the minimum, its eigenvalues (eig_host: [3N] or NULL),
 * sum of log over the non-zero modes, zero-mode count; OFLY_ERR_INVALID with the reference's message for a negative mode or an
 * unexpected zero-mode count (the outputs are still written). */
int ofly_mech_minimum(ofly_batch *b, const double *in_min_host, double zero_tol, double *log_prod_eigen, int *num_zero_modes,
                      int *count_frozen, double *eig_host);
/* do_adj_min + the numeric half of saddle_2_mech (find.cpp:1014-1090, 757-871) for n_saddles dimer results at once.
 * sp / axis: [S][3N] as Dimer::find_sp returned them; in_min: [3N]. rev / sp_out / fwd: [S][3N] drift-corrected pathway
 * (may be NULL). records: [S]. Uses b's R slots for energies / Hessians and a sibling batch for the relaxations. */
int ofly_mech_from_saddles(ofly_batch *b, const ofly_mech_options *opt, const ofly_lbfgs_options *min_opt, long n_saddles,
                           const double *sp_host, const double *axis_host, const double *in_min_host, double *rev_host,
                           double *sp_out_host, double *fwd_host, ofly_mech_record *records);
/* Master::perturb (find.cpp:1147-1214) for n_states start states at once. The batch's list is rebuilt from in_min at r_pert
 * (nl_pert, find.cpp:217-219). The caller draws the random numbers (the reference draws from a thread-local Xoshiro in
 * list order, which has no meaning here): stddev: [n] (the accepted N(stddev, stddev/3) draws), normals: [n][N][6] standard
 * normal draws per atom -- three for the displacement (scaled by stddev (1 - r / r_pert)), three for the axis; only atoms
 * within r_pert of `centre` use theirs. pos_out / axis_out: [n][3N]; axes are normalised. */
int ofly_master_perturb(ofly_batch *b, long n_states, const double *in_min_host, long centre, double r_pert, const double *stddev_host,
                        const double *normals_host, double *pos_out_host, double *axis_out_host);
/* The numeric part of one SKMC step (include/libfly/kinetic/skmc.hpp:297-309): E0 = energy(cell); relax the reconstructed
 * state; centroid_align(rel_recon, raw_recon); Ef = energy(rel_recon); dR = |rel_recon - raw_recon|. One call, every
 * intermediate on the device. rel_recon: [3N] or NULL. *min_failed follows LBFGS::minimise's bool. */
int ofly_skmc_relax_step(ofly_batch *b, const ofly_lbfgs_options *min_opt, const double *cell_host, const double *raw_recon_host,
                         double *rel_recon_host, double *E0, double *Ef, double *dR, int *min_failed);
/* find.cpp:287: sqrt(exp(log_prod_eigen_min - log_prod_eigen_sp) / (2 pi 1.6605390666050e-27)), in Hz. */
double ofly_mech_prefactor(double log_prod_eigen_min, double log_prod_eigen_sp);

/* Measurement aid (bench.py): FP64 FMA instructions per second this device sustains on register-only chains. */
int ofly_bench_fp64_fma(ofly_ctx *ctx, int iters, double *fma_per_s);

/* Plain device allocation helpers so a host program needs no CUDA headers of its own. */
int ofly_dev_alloc(ofly_ctx *ctx, size_t bytes, void **out_dev);
int ofly_dev_free(ofly_ctx *ctx, void *dev);
int ofly_dev_upload(ofly_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int ofly_dev_download(ofly_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* OPENFLY_B200_H */
