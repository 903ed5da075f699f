/* ptp.h -- C ABI of the B200-native PIC step for a Penning-Malmberg trapped plasma.
 *
 * Drop-in boundary for ONE hot path of Daniel32Duque/PIC-Trapped-Plasma:
 * PenningTrap::movePlasmas(dt) (reference Source/PenningTrap.cpp:352-363) and the
 * private phases below it. The reference has no FFI layer; its boundary is the
 * C++ class surface (Source/PenningTrap.hpp:54-98, Source/Plasma.hpp:139-197).
 * The host-side classes in pic-trapped-plasma_b200/host keep that surface and
 * call the functions below; every function names the reference member it
 * replaces. Plain pointers and sizes only; all host buffers are caller-owned,
 * all device memory is library-owned unless a function says "device pointer".
 *
 * Conventions
 *   - return 0 on success, a PTP_E* code otherwise; ptp_last_error() gives text.
 *   - grids are r-major, z fastest: node (r=j, z=k) -> (Nz+1)*j + k, G=(Nz+1)*Nr
 *     doubles (reference layout, Source/PenningTrap.hpp:51).
 *   - one host thread per trap; calls are synchronous w.r.t. returned host data.
 *   - there is NO CPU fallback: every entry point fails with PTP_ECUDA when no
 *     CUDA device is usable.
 */
#ifndef PTP_H
#define PTP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ptp_trap ptp_trap;     /* device twin of PenningTrap (Source/PenningTrap.hpp:54-98) */
typedef struct ptp_plasma ptp_plasma; /* device twin of Plasma      (Source/Plasma.hpp:139-197)   */

enum {
	PTP_OK = 0,
	PTP_EINVAL = 1,   /* bad argument (the host classes turn these into std::logic_error) */
	PTP_ECUDA = 2,    /* CUDA runtime error / no device */
	PTP_ENOMEM = 3,
	PTP_ECOMM = 4,    /* NCCL error or NCCL not loadable */
	PTP_ESTATE = 5    /* call made in the wrong state (e.g. step before upload) */
};

/* deposit accumulator (north_star: fp64 default, fixed-point = deterministic) */
enum { PTP_DEPOSIT_FP64 = 0, PTP_DEPOSIT_FIXED64 = 1 };
/* PTP_ARITH_FAST: reciprocal multiplies for /hz, /mass (cell index still bit-exact).
 * PTP_ARITH_EXACT: IEEE divisions in the reference's expression order -> per-ring
 * z, v bit-identical to the reference for identical node fields. */
enum { PTP_ARITH_FAST = 0, PTP_ARITH_EXACT = 1 };
/* Poisson solver: direct separable (DCT-I in z + Thomas in r), red-black SOR (cross-check), or the direct solver with
 * the inverse DCT forced through the FFT kernel (needs a power-of-two Nz; chosen automatically for long rows). */
enum { PTP_SOLVER_DIRECT = 0, PTP_SOLVER_SOR = 1, PTP_SOLVER_DIRECT_FFT = 2 };

const char* ptp_last_error(void);
int ptp_version(void);
int ptp_device_count(void);

/* ---- trap ------------------------------------------------------------------------------- */

/* PenningTrap::PenningTrap + generateSparse + analyzePattern/factorize
 * (Source/PenningTrap.cpp:36-58, 94-162): builds the device operator for the
 * 5-point cylindrical stencil. hz = length/Nz and hr = radius/Nr are passed as
 * the host class computed them (Source/PenningTrap.cpp:51-52). `device` is the
 * CUDA ordinal. phi_trap starts at zero: call ptp_trap_set_wall next. */
int ptp_trap_create(ptp_trap** out, int Nz, int Nr, double hz, double hr, double length, double radius, int device);
int ptp_trap_destroy(ptp_trap* t);

/* PenningTrap::updateRHS + solveLaplace (Source/PenningTrap.cpp:163-203), also what
 * setPotential (Source/PenningTrap.cpp:313-317) triggers. vWall[k], k=0..Nz, is the
 * wall potential per axial node (the `boundary` values of :169-197; the host class
 * keeps that electrode->node loop). Sets RHS(last row) = -(hr^-2 + 1/(2(R-hr)hr))*vWall
 * and solves for phi_trap. */
int ptp_trap_set_wall(ptp_trap* t, const double* vWall);

/* Time-dependent electrode programmes (PenningTrap::setPotential before every step, as in
 * Diagnostics/D) Useless Boundary Test.txt:120-134; Source/PenningTrap.cpp:313-317 = updateRHS + solveLaplace each time).
 * phi_trap is linear in the wall potential and the wall of PenningTrap::updateRHS (:169-197) is linear in the electrode
 * potentials (the gap ramps of :184 included), so phi_trap = sum_i V_i phi_i with phi_i the Laplace solution for "electrode
 * i at 1 V, all others grounded". ptp_trap_set_wall_basis registers nBasis wall profiles walls[nBasis][Nz+1] (one solve
 * each, once); ptp_trap_set_wall_weights then replaces a setPotential call by one axpy kernel (no solve, no wall upload),
 * and ptp_trap_step_programme runs nSteps steps with weights[s][nBasis] applied before step s without returning to the
 * host in between. Equal to the per-step solve up to the re-association of the sum (~1e-16 relative). */
int ptp_trap_set_wall_basis(ptp_trap* t, int nBasis, const double* walls);
int ptp_trap_set_wall_weights(ptp_trap* t, const double* weights);
int ptp_trap_step_programme(ptp_trap* t, double dt, int nSteps, const double* weights);

/* solver.solve(b) (Source/PenningTrap.cpp:202, Source/Plasma.cpp:98,389,412): phi = A^-1 rhs, host buffers of G. */
int ptp_trap_solve(ptp_trap* t, const double* rhs, double* phi);
/* y = A x with the assembled operator (coefficients * vector, Source/PenningTrap.cpp:250). */
int ptp_trap_apply(ptp_trap* t, const double* x, double* y);

int ptp_trap_get_phi(ptp_trap* t, double* phi);        /* potentialsVector (Source/PenningTrap.hpp:51) */
int ptp_trap_set_phi(ptp_trap* t, const double* phi);  /* parity hook: inject phi_trap */
/* PenningTrap::getEField(int,int) on every node (Source/PenningTrap.cpp:208-236), for the current potentials. */
int ptp_trap_get_enodes(ptp_trap* t, double* eNodes);

/* PenningTrap::movePlasmas(dt) x nSteps (Source/PenningTrap.cpp:352-363): every plasma pushed with the
 * pre-step field, then every plasma deposited and solved. */
int ptp_trap_step(ptp_trap* t, double dt, int nSteps);
/* The two halves of a step, separately (parity and timing): Plasma::moveRings + Plasma::updateRHS of all
 * plasmas (Source/Plasma.cpp:100-120, 77-94) ... */
int ptp_trap_push_deposit(ptp_trap* t, double dt);
/* ... and solver.solve of every plasma's RHS plus the node field (Source/Plasma.cpp:98; PenningTrap.cpp:208-236). */
int ptp_trap_solve_fields(ptp_trap* t);
/* Block until all queued work of this trap has finished. */
int ptp_trap_sync(ptp_trap* t);
/* Device time in milliseconds of the last ptp_trap_step call: [0]=whole call; summed over its steps: [1]=push+deposit kernels,
 * [2]=all-reduce, [3]=solve + node field (CUDA events on the trap's stream). The per-phase entries are zero unless
 * ptp_trap_set_phase_events(t, 1) was called and the steps were stream-launched (not replayed as a graph): events between
 * the kernels of a step cost the programmatic-launch overlap, so they are recorded on request only. */
int ptp_trap_last_times(ptp_trap* t, double* ms4);
int ptp_trap_set_phase_events(ptp_trap* t, int on);
/* Number of kernels the last ptp_trap_step / push_deposit / solve_fields call launched. */
int64_t ptp_trap_last_launches(ptp_trap* t);

/* Maintenance (K5): per-row counting sort of every plasma's rings by axial cell and compaction of lost rings. */
int ptp_trap_sort(ptp_trap* t);
/* Re-sort policy inside ptp_trap_step / ptp_trap_step_programme. interval > 0: every `interval` steps; 0: never;
 * -1 (default): adaptive - the push kernel counts the deposits that missed its thread-private cell window (rings that
 * drifted out of the cell range their segment was planned for: long plasmas on fine grids), the counters are read every
 * PTP_SORT_CHECK_STEPS steps (environment, default 16) and a species is re-sorted when its miss rate has risen by more
 * than PTP_SORT_FAR_FRACTION (default 5e-5 per ring-step) over the rate measured right after its last load or sort. The reference keeps its rings in one std::vector in load order
 * (Source/Plasma.hpp:46); ring identities survive a sort (ptp_plasma_download returns the upload index of every ring).
 * CUDA-graph replay (ptp_trap_set_graph) runs without the adaptive check. */
int ptp_trap_set_sort_interval(ptp_trap* t, int interval);
/* Number of per-species re-sorts either policy has triggered since the trap was created. */
int64_t ptp_trap_sorts_done(ptp_trap* t);
/* Hot species. Plasma::moveRings (Source/Plasma.cpp:100-120) is indifferent to the order of its rings; the default push kernel
 * is not - it needs the rings of a segment within a 44-cell window, which re-sorts maintain. Rings that cross the whole plasma
 * within a few dozen steps (electrons on a fine grid: more than a cell per step) cannot be kept ordered; for them the push
 * kernel has a second form (per-warp bins over one window of up to ~1400 cells, no re-sorts). mode 1: always that form,
 * 0: never, -1 (default; PTP_SCATTER overrides the default for new species): the adaptive re-sort policy switches a species
 * over when a third re-sort in a row is due less than PTP_HOT_SORT_STEPS (default 64) steps after the one before. Same results: identical bits in
 * fixed-point deposit mode, rounding level in fp64 mode; positions and speeds do not depend on the form at all. */
int ptp_plasma_set_hot(ptp_plasma* p, int mode);
/* 1 while the per-warp-bin form of the push kernel is in use for this species (valid after the first step or deposit). */
int ptp_plasma_is_hot(ptp_plasma* p);

int ptp_trap_set_deposit_mode(ptp_trap* t, int mode);
int ptp_trap_set_arith_mode(ptp_trap* t, int mode);
int ptp_trap_set_solver(ptp_trap* t, int solver, double sorTolerance, int sorMaxIterations);
/* Replay each step (or pair of steps in peer-memory mode) as a CUDA graph inside ptp_trap_step: removes the per-kernel
 * launch cost that dominates small configurations (the reference's default 8 k-ring case). Per-phase times are not
 * recorded in this mode ([1..3] of ptp_trap_last_times read 0). */
int ptp_trap_set_graph(ptp_trap* t, int on);
/* Tuning of the push kernel: threads per CTA (256/512), cells of the thread-private deposit window, CTAs
 * (0 = one per SM), rings per thread and tile (4/8). 0 (ctas: -1) keeps the current value. */
int ptp_trap_set_tuning(ptp_trap* t, int threads, int window, int ctas, int ringsPerThread);

/* ---- multi-GPU (one process per GPU; rings sharded, rho all-reduced, solve replicated) --- */

/* 128-byte NCCL unique id; rank 0 creates it, the launcher broadcasts it (torch.distributed / MPI / file). */
int ptp_comm_unique_id(void* id128);
int ptp_trap_comm_init(ptp_trap* t, const void* id128, int nRanks, int rank);
/* Exchange of the deposit grids between the ranks (needs ptp_trap_comm_init):
 *   0 = NCCL all-reduce;
 *   1 = peer memory, fused: the push kernel's flush adds into every rank's grid over NVLink (system-scope atomics) and a
 *       flag barrier replaces the collective;
 *   3 = peer memory, gather: one small kernel per step pushes this rank's populated rows into a slot of every rank's gather
 *       area with plain stores, flags, and sums the slots in rank order - no remote atomics, every rank holds bitwise the
 *       same sums also in fp64 mode;
 *   2 = choose by grid size, as measured: fused up to 2^20 nodes, gather above.
 * Multi-rank callers must issue the same sequence of calls on every rank (same species created in the same order): the
 * peer mappings are exchanged collectively whenever the grids had to be reallocated. */
int ptp_trap_set_allreduce(ptp_trap* t, int kind);

/* ---- plasma ------------------------------------------------------------------------------- */

/* Plasma::Plasma (Source/Plasma.cpp:68-72): registers with the trap; registration order is the
 * summation order of the species' potentials in the node field (Source/PenningTrap.cpp:228-232). */
int ptp_plasma_create(ptp_trap* t, ptp_plasma** out, double mass, double charge);
int ptp_plasma_destroy(ptp_plasma* p);

/* Replace the rings (what both loaders end with, Source/Plasma.cpp:504-526 / 598-620): n rings
 * {r[i] in [0,Nr), z[i], v[i]} in any order; ring i keeps id i. macroChargeDensity as in
 * Source/Plasma.cpp:494. Does not deposit or solve. */
int ptp_plasma_upload(ptp_plasma* p, int64_t n, const int32_t* r, const double* z, const double* v, double macroChargeDensity);
/* Ring placement of Plasma::loadDensityFile / Plasma::loadProfile (Source/Plasma.cpp:464-526 = :558-620) on the device:
 * density[G] (host) is the expected charge density; the cumulative charge per row, chargeMacro (:492) and the rings per
 * row (:496,499) are evaluated on the host in the reference's serial order, the quantile inversion (:512-523, positions
 * bit-identical) and the Maxwellian speeds run on the device. The speeds reproduce the deviate stream of the reference's
 * freshly seeded std::default_random_engine + std::normal_distribution (:508-509, libstdc++: minstd_rand0 and the polar
 * method) ring for ring; values agree to the last bits of log(). Shard `shard` of `nShards` keeps rings i = shard
 * (mod nShards) of every row (the multi-GPU partition); ring ids count the shard's rings in (row, i) order.
 * chargeMacro / nAtRow[Nr] (rings per row over all shards) / nLoaded (rings of this shard) may be NULL.
 * Does not deposit or solve (call ptp_plasma_deposit_solve, as both loaders do at :528,622). */
int ptp_plasma_load_density(ptp_plasma* p, const double* density, double temperature, int64_t numMacro, int shard, int nShards,
	double* chargeMacro, int64_t* nAtRow, int64_t* nLoaded);
/* Plasma::solvePoisson (Source/Plasma.cpp:95-99): deposit this plasma's RHS and solve its self potential. */
int ptp_plasma_deposit_solve(ptp_plasma* p);
/* Plasma::updateRHS alone (Source/Plasma.cpp:77-94). */
int ptp_plasma_deposit(ptp_plasma* p);

int ptp_plasma_count(ptp_plasma* p, int64_t* nAlive);   /* Plasma::getNumMacro (Source/Plasma.cpp:147-150) */
/* Live rings, row-bucketed order; buffers of at least ptp_plasma_count entries; id = index at upload. Any pointer may be NULL. */
int ptp_plasma_download(ptp_plasma* p, int32_t* r, double* z, double* v, int64_t* id);
/* The integer keys of the step for the live rings, same order as ptp_plasma_download:
 * k = (int)floor(z/hz), idx = (Nz+1)*r + k (Source/Plasma.cpp:87-88). */
int ptp_plasma_cell_index(ptp_plasma* p, int32_t* k, int32_t* idx);

int ptp_plasma_get_rhs(ptp_plasma* p, double* rhs);                 /* Plasma::RHS (Source/Plasma.hpp:53) */
int ptp_plasma_get_self_potential(ptp_plasma* p, double* phi);      /* Plasma::selfPotential (:52) */
int ptp_plasma_set_self_potential(ptp_plasma* p, const double* phi);/* parity hook / loaders */

/* Diagnostics reductions on device ("next" row 2 of SURVEY 8f):
 * getPotentialEnergy (Source/Plasma.cpp:244-252), sum of ring mass*v^2 terms for getTemperature (:212-228),
 * getNumMacroCentralWell (:151-162; limits = per-row [left,right] node indices, Source/PenningTrap.cpp:63-90). */
int ptp_plasma_potential_energy(ptp_plasma* p, double chargeMacro, double* pe);
int ptp_plasma_count_central_well(ptp_plasma* p, const int32_t* limitLeft, const int32_t* limitRight, int64_t* n);
/* Plasma::getTemperature / getAverageTemperature / getstdDeviation (Source/Plasma.cpp:163-228) without histories on the host:
 * over the live rings, *sumW = sum of w (w = 1 on the axis, 8 r elsewhere: ring mass = w massMacro, :169-170) and
 * *sumWS2 = sum of w speed^2 with speed = mean of the ring's speeds at the last two save points (:180, :224) - the present
 * speed when there is no earlier save point (*paired = 0). T = mass * sumWS2 / (KB * sumW). markSavePoint != 0 makes the
 * present speeds the save point (what PenningTrap::saveStates does, Source/PenningTrap.cpp:364-385); the saved speeds
 * follow their rings through re-sorts. Rings lost between two save points drop out of later sums, as in the reference. */
int ptp_plasma_kinetic_sums(ptp_plasma* p, int markSavePoint, double* sumW, double* sumWS2, int* paired);
/* Plasma::saveState(int indexR) (Source/Plasma.cpp:338-346): the live rings of ONE radial row - a contiguous slice of the
 * row-bucketed storage, so only that slice crosses PCIe. z, v, id (any may be NULL) hold nMax entries; *n = rings in the row. */
int ptp_plasma_download_row(ptp_plasma* p, int row, int64_t nMax, double* z, double* v, int64_t* id, int64_t* n);
/* Loss log of Plasma::moveRings (Source/Plasma.cpp:108-118): the reference removes a lost ring at once by swapping it with
 * the last one, step by step, which fixes the row order of its history files. The push kernel logs (ring id, step tag) for
 * every ring that leaves the trap (step tag = pushes of this species since its load); entries [first, first + nMax) are
 * copied out in the order they were logged (steps ascending), *total = entries logged so far, *overflowed = 1 when more
 * rings were lost than the log holds (65536; the host classes then fall back to one sweep per refresh). */
int ptp_plasma_loss_log(ptp_plasma* p, int64_t first, int64_t nMax, int64_t* ids, int64_t* steps, int64_t* total, int* overflowed);

#ifdef __cplusplus
}
#endif
#endif /* PTP_H */
