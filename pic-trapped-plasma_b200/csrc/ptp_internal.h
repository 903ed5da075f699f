// Internal declarations shared by the .cu files of libptp_b200.so. Not part of the ABI (see include/ptp.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ptp.h"

#define PTP_VERSION 100

// Row buckets are padded to this many ring slots so that every tile of the push kernel lies inside one
// radial row and double2 accesses stay 16-byte aligned (8 rings/thread x 512 threads max).
#define PTP_ROW_ALIGN 4096
// Rows per block of the radial product tables of the large-grid solver (multiple of 8).
#define PTP_THOMAS_BLOCK 32

// One contiguous run of ring slots of ONE radial row, owned by one CTA of the push kernel.
struct PtpSegment {
	int row;
	int pad;
	long long begin, end; // slot range, multiples of the tile size
};

struct ptp_plasma {
	ptp_trap* trap = nullptr;
	int index = -1;              // position in trap->plasmas = summation order of the node field
	double mass = 0, charge = 0, macroChargeDensity = 0;
	int64_t nUploaded = 0;       // rings given at upload
	int64_t nAlive = 0;          // host copy, refreshed from the device loss counter
	long long cap = 0;           // ring slots allocated (sum of padded row buckets)
	double* z = nullptr;         // [cap] axial position, NaN = empty slot / lost ring
	double* v = nullptr;         // [cap] axial speed at t - dt/2
	long long* id = nullptr;     // [cap] index of the ring at upload (moves only in sort/compaction)
	double* zAlt = nullptr;      // sort ping-pong buffers (allocated on first sort)
	double* vAlt = nullptr;
	long long* idAlt = nullptr;
	double farBaseline = -1.0;    // out-of-window deposits per ring-step right after the last load / sort (what a sort cannot remove:
	                              // the sparse tails of a row, whose tiles are wider than the window); < 0: not measured yet
	std::vector<long long> altDirty; // [Nr] slots at the start of each bucket of the alternate buffers that do not hold the empty-slot pattern
	void* sortScratch = nullptr;  // counters, cursors and chunk table of the sort (kept between sorts)
	size_t sortScratchBytes = 0;
	std::vector<long long> rowOff;   // [Nr+1] slot offset of each row bucket (host)
	std::vector<long long> rowLive;  // [Nr] slots of the bucket that may hold live rings (prefix of the bucket)
	long long* dRowOff = nullptr;
	std::vector<PtpSegment> segs;
	std::vector<int> ctaSegBegin;    // [nCta+1]
	PtpSegment* dSegs = nullptr;
	int* dCtaSegBegin = nullptr;
	int4* dSegBounds = nullptr;      // per segment: min / max / mean axial cell of its live rings (x, y, z)
	size_t segCap = 0, ctaCap = 0;   // allocated entries of dSegs / dSegBounds and of dCtaSegBegin (grow-only: re-plans are frequent)
	void* planScratch = nullptr;     // tile list, tile bounds and counters of ptp_tile_bounds (grow-only)
	size_t planScratchBytes = 0;
	int nCta = 0;
	unsigned long long* dLost = nullptr; // [2] device counters: rings lost since upload; deposits outside the private window since the last check
	unsigned long long* dLossLog = nullptr; // loss log of the push kernel (header of 4 words + (id, step tag) pairs), reset by every (re)load
	long long lossCap = 0;               // entries the log can hold
	double* vSaved = nullptr;            // [cap] speed of every ring at the last save point (ptp_plasma_kinetic_sums), allocated on first use
	double* vSavedAlt = nullptr;         // sort ping-pong
	bool vSavedValid = false;
	bool boundsValid = false;
	double ctaShare = 1.0;       // share of the SMs this species' push gets when all species run in one launch (by live rings)
	// Hot species: rings that cross the whole plasma within a few dozen steps (electrons on a fine grid) cannot be kept ordered
	// by cell; they are pushed by the SCATTER variant of K1 (per-warp bins over one wide window, no re-sorts).
	int hot = -1;                // ptp_plasma_set_hot: -1 decided by the re-sort policy, 0 never, 1 always
	bool hotAuto = false;        // hot was set by the policy, not by the caller: a reload starts undecided again
	bool scatter = false;        // the variant in use (segment tables are planned for it)
	long long lastSortStep = -1; // trap step count at the last miss-triggered re-sort (-1: none since the load)
	int quickSorts = 0;          // re-sorts in a row that came less than hotSortSteps steps after the one before
	bool encValid = false;       // the touched-node range per row kept next to this species' deposit grid (written by the push kernel's
	                             // flush) describes the grid's present content
};

struct PtpComm; // ptp_comm.cu

struct ptp_trap {
	int device = 0;
	int Nz = 0, Nr = 0;
	long long G = 0;
	double hz = 0, hr = 0, length = 0, radius = 0;
	int smCount = 148;
	size_t smemMax = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	bool phaseEvents = false;        // ptp_trap_set_phase_events: record 4 events per step (they sit between the kernels of the
	                                 // programmatic-launch chain and cost the overlap, so they are off unless a caller asks for phase times)
	std::vector<cudaEvent_t> evPool; // 4 events per step of the last ptp_trap_step call (phase timing)
	int evSteps = 0;
	double lastMs[4] = { 0, 0, 0, 0 };
	int64_t lastLaunches = 0;

	// operator / direct solver (ptp_solve.cu)
	double* solverConst = nullptr; // one allocation: dctInv | dctFwd | thInv | thCp (L2-persisting window)
	double* dctFwd = nullptr;    // [(Nz+1)^2]  FT[k][m] = (2/Nz) w_k w_m cos(pi k m / Nz)
	double* dctInv = nullptr;    // [(Nz+1)^2]  C[m][k]  = cos(pi m k / Nz)
	double* thInv = nullptr;     // [Nr][Nz+1]  1 / pivot of the r-tridiagonal of axial mode m
	double* thCp = nullptr;      // [Nr][Nz+1]  upper / pivot
	double* thR = nullptr;       // [Nr][Nz+1]  x_j / x_{j-1} above the outermost deposit row (factorisation from the wall)
	double* thQ = nullptr;       // [Nr][Nz+1]  1 / (pivot + upper * thR[j+1]): closes the downward sweep at that row
	double* thP = nullptr;       // [Nr][Nz+1]  product of thR over rows (block start .. j), blocks of PTP_THOMAS_BLOCK rows
	double* wideXb = nullptr;    // [species][blocks][Nz+1] value entering each block above the deposit (k_thomas_wide -> k_thomas_expand)
	int* wideJ = nullptr;        // [species] outermost deposit row of the last solve
	int wideCap = 0;
	double* thLower = nullptr;   // [Nr]        sub-diagonal of T_r
	double* stLower = nullptr;   // [Nr] stencil r-lower   (operator apply / SOR)
	double* stUpper = nullptr;   // [Nr] stencil r-upper
	double stDiag = 0, stHz2 = 0, wallFactor = 0;
	double2* fftTw = nullptr;    // [Nz] exp(-i pi j / Nz), only when Nz is a power of two (FFT inverse transform)
	int2* rowBounds = nullptr;   // [species x Nr] non-zero axial range of each deposit row (forward transform)
	int rowBoundsCap = 0;

	double* basisPhi = nullptr;  // [nBasis][G] Laplace solutions of the registered wall basis (electrode programmes)
	double* dWeights = nullptr;  // [weightsCap] weights of the programme's steps
	int nBasis = 0;
	size_t weightsCap = 0;
	double* phiTrap = nullptr;   // [G]
	double* eNodes = nullptr;    // [G]
	double* tmpA = nullptr;      // [G] scratch (host-RHS solves, wall RHS)
	double* tmpB = nullptr;      // [G]
	double* tmpSpec = nullptr;   // [G]

	// per-species grids, contiguous over species so that one all-reduce / one batched solve covers them
	int capS = 0;
	size_t spanDoubles = 0;      // one parity of rhoStore: capS*G accumulators + capS*Nr row bounds (8 bytes each)
	double* rhoStore = nullptr;  // one allocation: 2 parities x { [capS][G] deposit accumulators, [capS][Nr] touched node range
	                             // per row (two u32, maintained by the push kernel's flush) } + 64 barrier flags;
	                             // IPC-shared with the other ranks in peer-memory mode
	int rhoParity = 0;
	bool peerStale = true;       // rhoStore was (re)allocated: the peers' mappings must be exchanged again
	double* rhoAll = nullptr;    // = rhoStore + rhoParity*capS*G: [capS][G] current accumulators (double weights / int64 fixed point)
	double* phiSelfAll = nullptr;// [capS][G]
	double* specAll = nullptr;   // [capS][G] spectral workspace
	double* dScale = nullptr;    // [capS] rho -> RHS factor per species
	std::vector<ptp_plasma*> plasmas;

	int depositMode = PTP_DEPOSIT_FP64;
	int arithMode = PTP_ARITH_FAST;
	int solver = PTP_SOLVER_DIRECT;
	double sorTol = 1e-12;
	int sorMaxIter = 20000;
	int fixedBits = 40;
	int threads = 512, window = 44, ctas = 0, ringsPerThread = 4;
	int sortInterval = -1;           // > 0: re-sort every so many steps; 0: never; -1: when the push kernel reports too many out-of-window deposits
	int stepsSinceCheck = 0;         // adaptive mode: steps since the out-of-window counters were last read
	int nextCheckSteps = 4;          // adaptive mode: steps until the next read (short right after a load / sort: measures the baseline)
	int sortCheckSteps = 16;         // adaptive mode: steps between two reads of the counters (PTP_SORT_CHECK_STEPS)
	double sortFarFraction = 5e-5;   // adaptive mode: re-sort a species when more than this fraction of its deposits missed the window (PTP_SORT_FAR_FRACTION)
	// kernel variants, read from the environment once, when the trap is created
	int fftR16 = 1;                  // PTP_FFT_R16: rows of 4096 nodes go through the radix-16 inverse (0: radix-2 pass pairs)
	int fftFormRows = 1;             // PTP_FFT_FORM_ROWS: the radix-16 inverse forms the rows above the plasma itself (0: k_thomas_expand)
	int scatterPolicy = -1;          // PTP_SCATTER: default of ptp_plasma::hot for new species (-1 automatic, 0 never, 1 always)
	int hotSortSteps = 64;           // automatic policy: a species whose re-sorts keep coming less than this many steps apart is hot (PTP_HOT_SORT_STEPS)
	int planSlack = -1;              // rows whose rings span more cells than the deposit window are cut into segments that leave this many
	                                 // cells of the window free (room for the rings' drift until the next re-sort); -1: window / 2 (PTP_PLAN_SLACK)
	long long sortsDone = 0;         // re-sorts triggered by either policy (ptp_trap_sorts_done)
	long long stepCount = 0;
	bool eNodesValid = false;
	// The push reads the node field only in rows that hold rings, and rings never change their row: a step solves for the
	// populated rows only (lazyRows; PTP_FULL_SOLVE=1 turns that off). phiRows = leading rows of phiSelfAll / eNodes that are
	// current; the whole grids are produced on demand from the deposit grids (ptp_materialize_fields) when a getter asks.
	bool lazyRows = true;
	int phiRows = 0;
	bool usePdl = true;              // PTP_PDL=0: plain stream-ordered launches
	int clusterSolve = 1;            // PTP_CLUSTER_SOLVE=0: the step's solve through the two-kernel path even when the plasma occupies few rows
	int multiPush = 1;               // PTP_MULTI_PUSH=0: one push launch per species
	int invBulk = 1;                 // PTP_INV_BULK=0: the cp.async form of the dense inverse transform also for even row lengths

	PtpComm* comm = nullptr;
	int allreduceKind = 0;
	// Rings never change their radial row (posR is immutable, Source/Plasma.hpp:22-24), so deposits can only land in rows
	// below rowExtent = 1 + the outermost populated row over all species and ranks: the all-reduce and (on large grids) the
	// per-step clearing of the deposit grids are restricted to those rows.
	long long layoutEpoch = 0;       // bumped by every (re)load of rings
	long long extentEpoch = -1;      // layoutEpoch rowExtent was computed for
	int rowExtent = 0;
	long long cleanEpoch = -1;       // layoutEpoch for which the rows >= rowExtent of rhoStore are known to be zero
	long long peerCleanEpoch = -1;   // peer-memory mode: layoutEpoch for which both parities were cleared collectively (the invariant
	                                 // "the parity not in use is zero" then carries over from call to call)

	// CUDA-graph replay of the step (ptp_trap_set_graph): 1 on, 0 off, -1 automatic (small loads, where launch overhead counts)
	int useGraph = -1;
	long long cfgEpoch = 0;          // bumped by everything that changes what a step launches (uploads, modes, tuning, ...)
	long long graphMaxRings = 8000000;   // automatic policy: replay up to this many rings on this GPU (PTP_GRAPH_MAX_RINGS)
	cudaGraphExec_t graphExec[2] = { nullptr, nullptr };   // one step each: [p] takes the deposit grids to parity p
	long long graphCfg = -1;
	double graphDt = 0;
	int64_t graphLaunches = 0;
	int graphRows = 0;               // phiRows after a replayed step
};

// ---- error handling -------------------------------------------------------------------------
void ptp_set_error(const std::string& msg);
int ptp_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define PTP_CUDA(call)                                                                    \
	do {                                                                                  \
		cudaError_t e_ = (call);                                                          \
		if (e_ != cudaSuccess) return ptp_cuda_fail(e_, #call, __FILE__, __LINE__);        \
	} while (0)
#define PTP_TRY(call)                   \
	do {                                \
		int rc_ = (call);               \
		if (rc_ != PTP_OK) return rc_;  \
	} while (0)

// ---- programmatic dependent launch (sm_90+) ----------------------------------------------------
// The kernels of a step form a chain on one stream: K1 per species -> [peer barrier] -> forward transform + radial solves ->
// inverse transform + node field -> K1 of the next step. Launched with the programmatic-stream-serialization attribute, a
// kernel's CTAs are scheduled as soon as every CTA of its predecessor has executed griddepcontrol.launch_dependents (all
// kernels of the chain do so first thing), which takes the grid-launch latency and the predecessor's tail off the critical
// path; griddepcontrol.wait then blocks until the predecessor has completed and its writes are visible. Every kernel of
// the chain executes the wait before it touches anything an earlier kernel of the chain writes - which also keeps the
// completion order transitive. What a kernel does before its wait reads only tables that no kernel writes (solver
// constants, segment tables). Kernels launched without the attribute behave as before.
#ifdef __CUDACC__
__device__ __forceinline__ void ptp_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void ptp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t ptp_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args&&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid;
	cfg.blockDim = block;
	cfg.dynamicSmemBytes = smem;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

// ---- ptp_solve.cu ------------------------------------------------------------------------------
int ptp_solver_build(ptp_trap* t);
void ptp_solver_free(ptp_trap* t);
// phi[s] = A^-1 (scale[s] * rho[s]) for nS consecutive grids; rho is double weights or int64 fixed point.
// withField: nS covers ALL species (phi = phiSelfAll) and the node field is produced too (fused when possible).
// encBounds: per (species,row) touched node range as written by the push kernel (nullptr: scan rho for non-zeros).
// rowLimit: radial rows >= rowLimit of rho are known to be zero (deposit grids: no ring lives there); -1 = unknown.
// rowsWanted > 0: only the first rowsWanted radial rows of phi (and of the node field) are needed; *rowsDone = rows produced.
int ptp_solver_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* spec, double* phi, bool withField = false, const uint2* encBounds = nullptr, int rowLimit = -1, int rowsWanted = 0, int* rowsDone = nullptr);
int ptp_solver_apply(ptp_trap* t, const double* x, double* y);
// ptp_solve_wide.cu: the same direct solve organised for large grids
bool ptp_solver_fft_fits(const ptp_trap* t);
// expand = false: rows above the block of the outermost deposit row are left to the inverse transform (rowsFormed = false there)
int ptp_solver_forward_wide(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* spec, const uint2* encBounds, bool expand = true, int rowLimit = -1, int rowsOut = 0);
bool ptp_solver_inverse_forms_rows(const ptp_trap* t);
int ptp_solver_inverse_fft(ptp_trap* t, const double* spec, double* phi, int nS, bool withField, bool rowsFormed = true, int rowsOut = 0);
// ptp_solve_cluster.cu: forward transform + radial solves + inverse transform + node field of a step in one cluster kernel
bool ptp_solver_cluster_plan(const ptp_trap* t, int nS, int rowLimit, int rowsOut, int* PM, int* NC, int* KWc, int* CW, size_t* smem);
int ptp_solver_cluster_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* phi, const uint2* encBounds, int rowLimit, int rowsOut,
	int PM, int NC, int KWc, int CW, size_t smem);
int ptp_node_field(ptp_trap* t);
int ptp_wall_rhs(ptp_trap* t, const double* dWall, double* dRhs);
int ptp_sor_run(ptp_trap* t, const double* rho, bool rhoIsFixed, const double* dScale, int nS, double* phi);

// ---- ptp_push.cu ---------------------------------------------------------------------------------
int ptp_push_launch(ptp_trap* t, ptp_plasma* p, double dt, bool push);
int ptp_push_launch_multi(ptp_trap* t, ptp_plasma* const* ps, int n, double dt);   // K1 of up to 4 species in one launch
int ptp_bounds_launch(ptp_trap* t, ptp_plasma* p);
int ptp_tile_bounds(ptp_trap* t, ptp_plasma* p, const std::vector<PtpSegment>& tiles, std::vector<int2>& tileBounds, int64_t* nLive);
size_t ptp_push_smem_bytes(const ptp_trap* t, int threads, int window);
int ptp_push_scatter_window(const ptp_trap* t);                  // cells of the SCATTER variant's window (0: does not fit)
size_t ptp_push_scatter_smem_bytes(const ptp_trap* t);
bool ptp_push_scatter_usable(const ptp_trap* t);                 // default tuning (512 x 4) and a window of a useful size
int ptp_push_field_window(const ptp_trap* t);
int ptp_push_configure(ptp_trap* t);

// ---- ptp_particles.cu ----------------------------------------------------------------------------
int ptp_build_segments(ptp_trap* t, ptp_plasma* p);
int ptp_plasma_set_layout(ptp_plasma* p, const std::vector<long long>& count, int64_t n, double macroChargeDensity);
int ptp_sort_plasma(ptp_trap* t, ptp_plasma* p);

// ---- ptp_comm.cu ---------------------------------------------------------------------------------
int ptp_comm_allreduce(ptp_trap* t, void* buf, size_t count, bool isInt64);
int ptp_comm_sum_int(ptp_trap* t, int* value, int n = 1);        // collective element-wise sum (the caller keeps the sums below 2^31)
int ptp_comm_max_int(ptp_trap* t, int* value, int n = 1);        // collective element-wise max over the ranks (synchronises the stream)
int ptp_row_extent(ptp_trap* t, int* extent);                    // collective on the first call after a (re)load, cached afterwards
int ptp_layout_sync(ptp_trap* t);                                // row extent + fixed-point scale agreed between the ranks (same caching)
int ptp_materialize_fields(ptp_trap* t);                         // ptp_api.cu: whole-grid potentials and node field when only the populated rows are current
void ptp_comm_free(ptp_trap* t);
int ptp_comm_size(ptp_trap* t);
int ptp_comm_rank(ptp_trap* t);
// peer-memory mode (ptp_trap_set_allreduce(t, 1)): the push kernel's flush adds into every rank's grid over NVLink
bool ptp_peer_mode(ptp_trap* t);                                  // either peer-memory exchange
bool ptp_peer_fused(ptp_trap* t);                                 // kind 1: the push kernel's flush adds into every rank's grid
bool ptp_peer_gather(ptp_trap* t);                                // kind 3: all-gather of the ranks' grids by stores + local sum in rank order
int ptp_peer_exchange(ptp_trap* t);                               // the gather exchange of this step (one kernel)
int ptp_peer_prepare(ptp_trap* t);                               // collective: (re)map the peers' rhoStore
void ptp_peer_targets(ptp_trap* t, int parity, size_t offsetDoubles, void** out, int* n); // grid pointers of all ranks
int ptp_peer_barrier(ptp_trap* t);                                // all ranks' pushes of this epoch have landed
int ptp_solver_reserve(ptp_trap* t, int nS);                       // allocate what ptp_solver_run would allocate lazily
