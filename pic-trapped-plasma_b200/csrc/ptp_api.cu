// C ABI (include/ptp.h): handles, grids, and the orchestration of PenningTrap::movePlasmas
// (reference Source/PenningTrap.cpp:352-363) on one CUDA stream per trap.
#include "ptp_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

// phi_trap = sum_i w_i phi_i over the registered wall basis (electrode programmes): fixed summation order.
__global__ void k_combine_basis(const double* __restrict__ basisPhi, const double* __restrict__ weights, int nBasis, long long G, double* __restrict__ phiTrap)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= G) return;
	double s = __dmul_rn(weights[0], basisPhi[idx]);
	for (int i = 1; i < nBasis; ++i) s = __dadd_rn(s, __dmul_rn(weights[i], basisPhi[(size_t)i * G + idx]));
	phiTrap[idx] = s;
}

int combine_basis(ptp_trap* t, const double* dWeights)
{
	k_combine_basis<<<(unsigned)((t->G + 255) / 256), 256, 0, t->stream>>>(t->basisPhi, dWeights, t->nBasis, t->G, t->phiTrap);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_combine_basis launch", __FILE__, __LINE__);
	t->lastLaunches++;
	t->eNodesValid = false;
	return PTP_OK;
}

thread_local std::string g_error;
}

void ptp_set_error(const std::string& msg) { g_error = msg; }

int ptp_cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
	g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" + std::to_string(line) + ")";
	return e == cudaErrorMemoryAllocation ? PTP_ENOMEM : PTP_ECUDA;
}

namespace {

// (Re)size the per-species grids so that species s owns slice s of each [capS][G] array.
int ensure_species_capacity(ptp_trap* t, int need)
{
	if (need <= t->capS) return PTP_OK;
	int cap = t->capS ? t->capS : 2;
	while (cap < need) cap *= 2;
	const size_t bytes = (size_t)cap * t->G * sizeof(double), old = (size_t)t->capS * t->G * sizeof(double);
	double *rho = nullptr, *phi = nullptr, *spec = nullptr, *scale = nullptr;
	const size_t span = (size_t)cap * t->G + (size_t)cap * t->Nr;       // doubles per parity: grids + row bounds
	const size_t storeBytes = 2 * span * sizeof(double) + 64 * sizeof(unsigned long long);
	cudaError_t e = cudaMalloc(&rho, storeBytes);
	if (e == cudaSuccess) e = cudaMalloc(&phi, bytes);
	if (e == cudaSuccess) e = cudaMalloc(&spec, bytes);
	if (e == cudaSuccess) e = cudaMalloc(&scale, cap * sizeof(double));
	if (e == cudaSuccess) e = cudaMemset(rho, 0, storeBytes);
	if (e == cudaSuccess) e = cudaMemset(phi, 0, bytes);
	if (e == cudaSuccess) e = cudaMemset(spec, 0, bytes);
	if (e == cudaSuccess) e = cudaMemset(scale, 0, cap * sizeof(double));
	if (e == cudaSuccess && t->capS) {
		e = cudaMemcpy(rho, t->rhoAll, old, cudaMemcpyDeviceToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(phi, t->phiSelfAll, old, cudaMemcpyDeviceToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(scale, t->dScale, t->capS * sizeof(double), cudaMemcpyDeviceToDevice);
	}
	if (e != cudaSuccess) {                                      // nothing of the trap has been touched yet
		cudaFree(rho); cudaFree(phi); cudaFree(spec); cudaFree(scale);
		return ptp_cuda_fail(e, "ensure_species_capacity", __FILE__, __LINE__);
	}
	for (ptp_plasma* p : t->plasmas) p->encValid = false;       // the touched-node ranges were not carried over: the solver scans
	cudaFree(t->rhoStore); cudaFree(t->phiSelfAll); cudaFree(t->specAll); cudaFree(t->dScale);
	t->rhoStore = rho; t->rhoParity = 0; t->rhoAll = rho; t->peerStale = true; t->spanDoubles = span;
	++t->cfgEpoch; ++t->layoutEpoch;
	t->phiSelfAll = phi; t->specAll = spec; t->dScale = scale;
	t->capS = cap;
	return PTP_OK;
}

// rowsWanted > 0 (only with withField, i.e. all species): the leading rows the caller needs; t->phiRows is updated.
int solve_species(ptp_trap* t, int first, int count, bool withField = false, int rowsWanted = 0)
{
	const bool fixed = t->depositMode == PTP_DEPOSIT_FIXED64;
	const double* rho = t->rhoAll + (size_t)first * t->G;
	double* phi = t->phiSelfAll + (size_t)first * t->G;
	if (t->solver == PTP_SOLVER_SOR) {
		PTP_TRY(ptp_sor_run(t, rho, fixed, t->dScale + first, count, phi));
		if (withField) t->phiRows = t->Nr;
		return withField ? ptp_node_field(t) : PTP_OK;
	}
	// the touched node range per row comes from the push kernel's flush when it has seen every deposit of the grid (one GPU,
	// or peer-memory mode inside a step) and nothing has replaced the grids since; otherwise the grid is scanned
	bool trust = ptp_comm_size(t) == 1 || (withField && ptp_peer_mode(t));
	for (int s = first; s < first + count && trust; ++s) trust = t->plasmas[s]->encValid;
	const uint2* enc = trust ? reinterpret_cast<const uint2*>(t->rhoAll + (size_t)t->capS * t->G) + (size_t)first * t->Nr : nullptr;
	const int rowLimit = t->extentEpoch == t->layoutEpoch ? t->rowExtent : -1;
	int rowsDone = t->Nr;
	PTP_TRY(ptp_solver_run(t, rho, fixed, t->dScale + first, count, t->specAll + (size_t)first * t->G, phi, withField, enc, rowLimit, rowsWanted, &rowsDone));
	if (withField) t->phiRows = rowsDone;
	return PTP_OK;
}

// Several species: one push launch for all of them (default tuning only), the SMs shared out by live rings; segment tables are
// (re)planned where a species' share has changed. Idempotent - capture_step_graph runs it before the capture starts, so that
// nothing is planned (a synchronising device pass) inside the capture.
int plan_push(ptp_trap* t, bool* multiOut)
{
	const int nS = (int)t->plasmas.size();
	bool multi = t->multiPush && nS >= 2 && nS <= 4 && t->threads == 512 && t->ringsPerThread == 4;
	for (ptp_plasma* p : t->plasmas) multi = multi && p->cap > 0 && p->hot != 1;   // (a hot species has a launch of its own: other kernel variant)
	double total = 0;
	for (ptp_plasma* p : t->plasmas) total += (double)std::max<int64_t>(p->nAlive, 1);
	for (ptp_plasma* p : t->plasmas) {
		const double share = multi ? (double)std::max<int64_t>(p->nAlive, 1) / total : 1.0;
		if (std::abs(share - p->ctaShare) > 0.1 * std::max(share, p->ctaShare)) { p->ctaShare = share; p->boundsValid = false; }
	}
	for (ptp_plasma* p : t->plasmas) {
		if (!p->boundsValid) PTP_TRY(ptp_bounds_launch(t, p));
		multi = multi && p->nCta > 0;
	}
	*multiOut = multi;
	return PTP_OK;
}

// Plasma::moveRings + Plasma::updateRHS of every species (push with the pre-step field, deposit at the new position).
// The deposit grids are double-buffered by step parity: this step's sums go into the parity that the push kernels of the
// previous step zeroed (their populated rows and touched-node ranges; begin_steps zeroes everything after a (re)load), and
// this step's push kernels zero the other one - no memset between the kernels of a step, on one GPU as in peer-memory mode
// (where the other ranks may add into the grid of the NEXT step as soon as they have passed this step's barrier).
int push_deposit_all(ptp_trap* t, double dt)
{
	if (!t->eNodesValid) PTP_TRY(ptp_node_field(t));
	t->rhoParity ^= 1;
	t->rhoAll = t->rhoStore + (size_t)t->rhoParity * t->spanDoubles;
	bool multi = false;
	PTP_TRY(plan_push(t, &multi));
	const int nS = (int)t->plasmas.size();
	if (multi) {
		PTP_TRY(ptp_push_launch_multi(t, t->plasmas.data(), nS, dt));
		for (ptp_plasma* p : t->plasmas) p->encValid = true;
		return PTP_OK;
	}
	for (ptp_plasma* p : t->plasmas) {
		if (!p->boundsValid) PTP_TRY(ptp_bounds_launch(t, p));
		if (p->cap == 0 || p->nCta == 0) {                       // no push kernel for an empty species: its slice of the other parity by hand
			double* other = t->rhoStore + (size_t)(t->rhoParity ^ 1) * t->spanDoubles;
			PTP_CUDA(cudaMemsetAsync(other + (size_t)p->index * t->G, 0, (size_t)t->G * sizeof(double), t->stream));
			PTP_CUDA(cudaMemsetAsync(other + (size_t)t->capS * t->G + (size_t)p->index * t->Nr, 0, (size_t)t->Nr * sizeof(double), t->stream));
		}
		PTP_TRY(ptp_push_launch(t, p, dt, true));
		p->encValid = true;
	}
	return PTP_OK;
}

// The exchange step: either an NCCL all-reduce of the rank-local grids, or - in peer-memory mode, where the push kernel has
// already added every rank's deposits into every rank's grid - just the barrier that says all of them have landed.
int reduce_rho(ptp_trap* t)
{
	const int nS = (int)t->plasmas.size();
	if (ptp_peer_gather(t)) return ptp_peer_exchange(t);
	if (ptp_peer_fused(t)) return ptp_peer_barrier(t);
	if (ptp_comm_size(t) == 1) return PTP_OK;
	int extent = t->Nr;
	PTP_TRY(ptp_row_extent(t, &extent));
	if (extent == t->Nr) return ptp_comm_allreduce(t, t->rhoAll, (size_t)nS * t->G, t->depositMode == PTP_DEPOSIT_FIXED64);
	for (int s = 0; s < nS; ++s)                                 // rows >= extent are zero on every rank
		PTP_TRY(ptp_comm_allreduce(t, t->rhoAll + (size_t)s * t->G, (size_t)extent * (t->Nz + 1), t->depositMode == PTP_DEPOSIT_FIXED64));
	return PTP_OK;
}

// At the start of every stepping call (ptp_trap_step, ptp_trap_push_deposit, ptp_trap_step_programme). Cheap when nothing
// has changed since the last call - per-step callers (the host classes call ptp_trap_step(dt, 1) from movePlasmas) pay for
// two comparisons. After a (re)load or a (re)allocation: the ranks agree on row extent and fixed-point scale (a small
// collective), potentials are completed where rings were loaded into rows the last solve left out, peer mappings are
// exchanged, and both parities of the deposit grids are zeroed - in peer-memory mode with a barrier behind the zeroing, so
// that no rank adds into a grid that its owner has not cleared yet. From then on the step keeps "the parity not in use is
// zero" by itself.
int begin_steps(ptp_trap* t)
{
	PTP_TRY(ptp_layout_sync(t));
	if (t->phiRows < std::min(t->rowExtent, t->Nr)) PTP_TRY(ptp_materialize_fields(t));   // (reads the deposit grids: before they are zeroed)
	const bool peer = ptp_peer_mode(t);
	if (peer) PTP_TRY(ptp_peer_prepare(t));
	if (t->cleanEpoch == t->layoutEpoch && (!peer || t->peerCleanEpoch == t->layoutEpoch)) return PTP_OK;
	PTP_CUDA(cudaMemsetAsync(t->rhoStore, 0, 2 * t->spanDoubles * sizeof(double), t->stream));
	for (ptp_plasma* p : t->plasmas) p->encValid = false;
	t->cleanEpoch = t->layoutEpoch;
	if (!peer) return PTP_OK;
	t->peerCleanEpoch = t->layoutEpoch;
	return ptp_peer_fused(t) ? ptp_peer_barrier(t) : PTP_OK;   // (gather exchange: nobody else writes into this rank's grids)
}

int solve_all(ptp_trap* t)
{
	if (t->plasmas.empty()) { t->phiRows = t->Nr; return ptp_node_field(t); }
	const bool lazy = t->lazyRows && t->extentEpoch == t->layoutEpoch && t->rowExtent > 0;
	return solve_species(t, 0, (int)t->plasmas.size(), true, lazy ? t->rowExtent : 0);
}

} // namespace

// Whole-grid potentials and node field from the deposit grids as they stand (the sums of the last step, or of the last
// stand-alone deposit): what the reference holds in Plasma::selfPotential after every solvePoisson (Source/Plasma.cpp:98).
// The arithmetic per row is the same as in the step's solve (same fold row, same transforms), so the rows the push used do
// not change by a bit.
int ptp_materialize_fields(ptp_trap* t)
{
	if (t->phiRows >= t->Nr) return PTP_OK;
	if (t->plasmas.empty()) { t->phiRows = t->Nr; return PTP_OK; }
	return solve_species(t, 0, (int)t->plasmas.size(), true, 0);   // (no collective here: a getter may run on one rank only)
}

extern "C" {

const char* ptp_last_error(void) { return g_error.c_str(); }
int ptp_version(void) { return PTP_VERSION; }

int ptp_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int ptp_trap_create(ptp_trap** out, int Nz, int Nr, double hz, double hr, double length, double radius, int device)
{
	if (!out) { ptp_set_error("ptp_trap_create: null output"); return PTP_EINVAL; }
	*out = nullptr;
	if (Nz < 4 || Nr < 2 || !(hz > 0) || !(hr > 0) || !(length > 0) || !(radius > 0)) {
		ptp_set_error("ptp_trap_create: need Nz >= 4, Nr >= 2 and positive hz, hr, length, radius");
		return PTP_EINVAL;
	}
	int nDev = 0;
	cudaError_t e = cudaGetDeviceCount(&nDev);
	if (e != cudaSuccess || nDev == 0) {
		cudaGetLastError();
		ptp_set_error("no CUDA device available (this library has no CPU fallback)");
		return PTP_ECUDA;
	}
	if (device < 0 || device >= nDev) { ptp_set_error("ptp_trap_create: bad device ordinal"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(device));
	ptp_trap* t = new ptp_trap;
	t->device = device;
	t->Nz = Nz; t->Nr = Nr; t->G = (long long)(Nz + 1) * Nr;
	t->hz = hz; t->hr = hr; t->length = length; t->radius = radius;
	t->phiRows = Nr;
	auto init = [&]() -> int {
		cudaDeviceProp prop;
		PTP_CUDA(cudaGetDeviceProperties(&prop, device));
		t->smCount = prop.multiProcessorCount;
		t->smemMax = prop.sharedMemPerBlockOptin;
		if (const char* e = std::getenv("PTP_FFT_R16")) t->fftR16 = std::atoi(e);
		if (const char* e = std::getenv("PTP_FFT_FORM_ROWS")) t->fftFormRows = std::atoi(e);
		if (const char* e = std::getenv("PTP_SCATTER")) t->scatterPolicy = std::atoi(e);
		if (const char* e = std::getenv("PTP_HOT_SORT_STEPS")) t->hotSortSteps = std::max(1, std::atoi(e));
		if (const char* e = std::getenv("PTP_PLAN_SLACK")) t->planSlack = std::atoi(e);
		if (const char* e = std::getenv("PTP_SORT_CHECK_STEPS")) t->sortCheckSteps = std::max(1, std::atoi(e));
		if (const char* e = std::getenv("PTP_SORT_FAR_FRACTION")) t->sortFarFraction = std::max(0.0, std::atof(e));
		if (const char* e = std::getenv("PTP_FULL_SOLVE")) t->lazyRows = std::atoi(e) == 0;
		if (const char* e = std::getenv("PTP_PDL")) t->usePdl = std::atoi(e) != 0;
		if (const char* e = std::getenv("PTP_INV_BULK")) t->invBulk = std::atoi(e);
		if (const char* e = std::getenv("PTP_MULTI_PUSH")) t->multiPush = std::atoi(e);
		if (const char* e = std::getenv("PTP_CLUSTER_SOLVE")) t->clusterSolve = std::atoi(e);
		if (const char* e = std::getenv("PTP_GRAPH")) t->useGraph = std::atoi(e);
		if (const char* e = std::getenv("PTP_GRAPH_MAX_RINGS")) t->graphMaxRings = std::atoll(e);
		PTP_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
		for (auto& ev : t->ev) PTP_CUDA(cudaEventCreate(&ev));
		const size_t gb = (size_t)t->G * sizeof(double);
		PTP_CUDA(cudaMalloc(&t->phiTrap, gb));
		PTP_CUDA(cudaMalloc(&t->eNodes, gb));
		PTP_CUDA(cudaMalloc(&t->tmpA, gb));
		PTP_CUDA(cudaMalloc(&t->tmpB, gb));
		PTP_CUDA(cudaMalloc(&t->tmpSpec, gb));
		PTP_CUDA(cudaMemset(t->phiTrap, 0, gb));
		PTP_CUDA(cudaMemset(t->eNodes, 0, gb));
		PTP_TRY(ptp_solver_build(t));
		return ensure_species_capacity(t, 2);
	};
	const int rc = init();
	if (rc != PTP_OK) {                                          // one cleanup path: whatever was created so far is released
		const std::string why = g_error;
		ptp_trap_destroy(t);
		g_error = why;
		return rc;
	}
	*out = t;
	return PTP_OK;
}

int ptp_trap_destroy(ptp_trap* t)
{
	if (!t) return PTP_OK;
	cudaSetDevice(t->device);
	if (t->stream) cudaStreamSynchronize(t->stream);
	while (!t->plasmas.empty()) ptp_plasma_destroy(t->plasmas.back());
	ptp_comm_free(t);
	ptp_solver_free(t);
	cudaFree(t->phiTrap); cudaFree(t->eNodes); cudaFree(t->tmpA); cudaFree(t->tmpB); cudaFree(t->tmpSpec);
	cudaFree(t->rhoStore); cudaFree(t->phiSelfAll); cudaFree(t->specAll); cudaFree(t->dScale);
	for (auto& ev : t->ev) if (ev) cudaEventDestroy(ev);
	for (auto& ev : t->evPool) cudaEventDestroy(ev);
	for (auto& g : t->graphExec) if (g) cudaGraphExecDestroy(g);
	cudaFree(t->basisPhi); cudaFree(t->dWeights);
	if (t->stream) cudaStreamDestroy(t->stream);
	delete t;
	return PTP_OK;
}

int ptp_trap_solve(ptp_trap* t, const double* rhs, double* phi)
{
	if (!t || !rhs || !phi) { ptp_set_error("ptp_trap_solve: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	const size_t gb = (size_t)t->G * sizeof(double);
	PTP_CUDA(cudaMemcpyAsync(t->tmpA, rhs, gb, cudaMemcpyHostToDevice, t->stream));
	if (t->solver == PTP_SOLVER_SOR) {
		PTP_CUDA(cudaMemsetAsync(t->tmpB, 0, gb, t->stream));
		PTP_TRY(ptp_sor_run(t, t->tmpA, false, nullptr, 1, t->tmpB));
	}
	else PTP_TRY(ptp_solver_run(t, t->tmpA, false, nullptr, 1, t->tmpSpec, t->tmpB));
	PTP_CUDA(cudaMemcpyAsync(phi, t->tmpB, gb, cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_apply(ptp_trap* t, const double* x, double* y)
{
	if (!t || !x || !y) { ptp_set_error("ptp_trap_apply: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	const size_t gb = (size_t)t->G * sizeof(double);
	PTP_CUDA(cudaMemcpyAsync(t->tmpA, x, gb, cudaMemcpyHostToDevice, t->stream));
	PTP_TRY(ptp_solver_apply(t, t->tmpA, t->tmpB));
	PTP_CUDA(cudaMemcpyAsync(y, t->tmpB, gb, cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_set_wall(ptp_trap* t, const double* vWall)
{
	if (!t || !vWall) { ptp_set_error("ptp_trap_set_wall: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	PTP_CUDA(cudaMemcpyAsync(t->tmpSpec, vWall, (size_t)(t->Nz + 1) * sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_TRY(ptp_wall_rhs(t, t->tmpSpec, t->tmpA));
	if (t->solver == PTP_SOLVER_SOR) PTP_TRY(ptp_sor_run(t, t->tmpA, false, nullptr, 1, t->phiTrap));
	else PTP_TRY(ptp_solver_run(t, t->tmpA, false, nullptr, 1, t->tmpSpec, t->phiTrap));
	t->eNodesValid = false;
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_set_wall_basis(ptp_trap* t, int nBasis, const double* walls)
{
	if (!t || nBasis < 1 || nBasis > 256 || !walls) { ptp_set_error("ptp_trap_set_wall_basis: bad arguments"); return PTP_EINVAL; }
	if (t->solver == PTP_SOLVER_SOR) { ptp_set_error("ptp_trap_set_wall_basis: needs the direct solver"); return PTP_ESTATE; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	const int n1 = t->Nz + 1;
	cudaFree(t->basisPhi); cudaFree(t->dWeights);
	t->basisPhi = nullptr; t->dWeights = nullptr; t->nBasis = 0; t->weightsCap = 0;
	PTP_CUDA(cudaMalloc(&t->basisPhi, (size_t)nBasis * t->G * sizeof(double)));
	for (int i = 0; i < nBasis; ++i) {                        // one Laplace solve per basis wall, as ptp_trap_set_wall does
		PTP_CUDA(cudaMemcpyAsync(t->tmpSpec, walls + (size_t)i * n1, (size_t)n1 * sizeof(double), cudaMemcpyHostToDevice, t->stream));
		PTP_TRY(ptp_wall_rhs(t, t->tmpSpec, t->tmpA));
		PTP_TRY(ptp_solver_run(t, t->tmpA, false, nullptr, 1, t->tmpSpec, t->basisPhi + (size_t)i * t->G));
	}
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	t->nBasis = nBasis;
	return PTP_OK;
}

namespace {
int upload_weights(ptp_trap* t, const double* weights, size_t count)
{
	if (t->weightsCap < count) {
		cudaFree(t->dWeights);
		t->dWeights = nullptr;
		PTP_CUDA(cudaMalloc(&t->dWeights, count * sizeof(double)));
		t->weightsCap = count;
	}
	PTP_CUDA(cudaMemcpyAsync(t->dWeights, weights, count * sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));               // the caller's buffer may be reused right away
	return PTP_OK;
}
} // namespace

int ptp_trap_set_wall_weights(ptp_trap* t, const double* weights)
{
	if (!t || !weights) { ptp_set_error("ptp_trap_set_wall_weights: null argument"); return PTP_EINVAL; }
	if (t->nBasis < 1) { ptp_set_error("ptp_trap_set_wall_weights: call ptp_trap_set_wall_basis first"); return PTP_ESTATE; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	PTP_TRY(upload_weights(t, weights, (size_t)t->nBasis));
	return combine_basis(t, t->dWeights);
}

int ptp_trap_get_phi(ptp_trap* t, double* phi)
{
	if (!t || !phi) { ptp_set_error("ptp_trap_get_phi: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_CUDA(cudaMemcpyAsync(phi, t->phiTrap, (size_t)t->G * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_set_phi(ptp_trap* t, const double* phi)
{
	if (!t || !phi) { ptp_set_error("ptp_trap_set_phi: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_CUDA(cudaMemcpyAsync(t->phiTrap, phi, (size_t)t->G * sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	t->eNodesValid = false;
	return PTP_OK;
}

int ptp_trap_get_enodes(ptp_trap* t, double* eNodes)
{
	if (!t || !eNodes) { ptp_set_error("ptp_trap_get_enodes: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_TRY(ptp_materialize_fields(t));
	if (!t->eNodesValid) PTP_TRY(ptp_node_field(t));
	PTP_CUDA(cudaMemcpyAsync(eNodes, t->eNodes, (size_t)t->G * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_push_deposit(ptp_trap* t, double dt)
{
	if (!t) { ptp_set_error("ptp_trap_push_deposit: null trap"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	PTP_TRY(begin_steps(t));
	PTP_TRY(push_deposit_all(t, dt));
	PTP_TRY(reduce_rho(t));
	return PTP_OK;
}

int ptp_trap_solve_fields(ptp_trap* t)
{
	if (!t) { ptp_set_error("ptp_trap_solve_fields: null trap"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	return solve_all(t);
}

namespace {

int one_step(ptp_trap* t, double dt, cudaEvent_t* e)
{
	if (e) PTP_CUDA(cudaEventRecord(e[0], t->stream));
	PTP_TRY(push_deposit_all(t, dt));
	if (e) PTP_CUDA(cudaEventRecord(e[1], t->stream));
	PTP_TRY(reduce_rho(t));
	if (e) PTP_CUDA(cudaEventRecord(e[2], t->stream));
	PTP_TRY(solve_all(t));
	if (e) PTP_CUDA(cudaEventRecord(e[3], t->stream));
	++t->stepCount;
	return PTP_OK;
}

// K5 policy after a step. Fixed interval: every sortInterval steps. Adaptive (sortInterval < 0): the push kernel counts the
// deposits that missed the thread-private window (rings that drifted away from the cell range their segment was planned
// for - long plasmas on fine grids); every sortCheckSteps steps the counters are read back, and a species whose miss rate
// has risen by more than sortFarFraction (and by more than half) over the rate measured right after its last load / sort
// is re-sorted by axial cell, which also re-plans its segments. Measured on the 4096 x 1024 grid: 1e-4 misses per
// ring-step already cost the push kernel 18 % (same-node global atomics), 5e-3 make it 4.6 x slower.
int maintain_order(ptp_trap* t)
{
	if (t->sortInterval > 0) {
		if (t->stepCount % t->sortInterval == 0)
			for (ptp_plasma* p : t->plasmas) { PTP_TRY(ptp_sort_plasma(t, p)); ++t->sortsDone; }
		return PTP_OK;
	}
	if (t->sortInterval == 0 || ++t->stepsSinceCheck < std::min(t->nextCheckSteps, t->sortCheckSteps)) return PTP_OK;
	const int steps = t->stepsSinceCheck;
	t->stepsSinceCheck = 0;
	t->nextCheckSteps = t->sortCheckSteps;
	const size_t nS = t->plasmas.size();
	std::vector<unsigned long long> h(2 * nS, 0ULL);
	for (size_t s = 0; s < nS; ++s)
		PTP_CUDA(cudaMemcpyAsync(&h[2 * s], t->plasmas[s]->dLost, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	for (size_t s = 0; s < nS; ++s) {
		ptp_plasma* p = t->plasmas[s];
		const unsigned long long far = h[2 * s + 1];
		const double alive = (double)(p->nUploaded - (int64_t)h[2 * s]);
		const double rate = alive > 0 ? (double)far / (alive * steps) : 0.0;
		bool sort;
		if (p->farBaseline < 0) {                               // first reading after a load / sort: what is left is not the drift's doing,
			sort = rate > 0.05;                                 // unless the rings came in unordered
			if (!sort) p->farBaseline = rate;
		}
		else sort = rate > p->farBaseline + std::max(t->sortFarFraction, 0.5 * p->farBaseline);
		if (p->scatter) sort = false;                           // (the wide window has no order to restore; misses there are rings beyond it)
		const bool quick = sort && p->lastSortStep >= 0 && t->stepCount - p->lastSortStep < t->hotSortSteps;
		if (quick && p->quickSorts >= 1 && p->hot < 0 && ptp_push_scatter_usable(t)) {
			// the third re-sort in a row that is due less than hotSortSteps steps after the one before: the rings of this species
			// cross the plasma faster than sorting can follow (each sort costs about ten steps). From here on the per-warp-bin
			// form of K1 pushes it (ptp_plasma_set_hot).
			p->hot = 1;
			p->hotAuto = true;
			PTP_TRY(ptp_build_segments(t, p));
			PTP_CUDA(cudaMemsetAsync(p->dLost + 1, 0, sizeof(unsigned long long), t->stream));
			sort = false;
		}
		else if (sort) {
			PTP_TRY(ptp_sort_plasma(t, p));                     // clears the counters and the baseline
			++t->sortsDone;
			p->quickSorts = quick ? p->quickSorts + 1 : 0;
			p->lastSortStep = t->stepCount;
			t->nextCheckSteps = 4;
		}
		else if (far) PTP_CUDA(cudaMemsetAsync(p->dLost + 1, 0, sizeof(unsigned long long), t->stream));
	}
	return PTP_OK;
}

void drop_graph(ptp_trap* t)
{
	for (auto& g : t->graphExec) {
		if (g) cudaGraphExecDestroy(g);
		g = nullptr;
	}
	t->graphCfg = -1;
}

// Capture ONE step into a graph: the step that takes the deposit grids from parity (target ^ 1) to parity `target` - two
// graphs, one per parity, so that any number of steps (also the one-step calls of PenningTrap::movePlasmas) can be
// replayed. Everything a step would allocate or synchronise on lazily is settled before the capture starts.
int capture_step_graph(ptp_trap* t, double dt, int target)
{
	{ bool multi; PTP_TRY(plan_push(t, &multi)); }
	if (!t->eNodesValid) PTP_TRY(ptp_node_field(t));
	PTP_TRY(ptp_solver_reserve(t, (int)t->plasmas.size()));
	const int parity0 = t->rhoParity;
	const long long steps0 = t->stepCount;
	const int64_t launches0 = t->lastLaunches;
	const int rows0 = t->phiRows;
	PTP_CUDA(cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
	const int rc = one_step(t, dt, nullptr);
	cudaGraph_t graph = nullptr;
	cudaError_t e = cudaStreamEndCapture(t->stream, &graph);
	t->graphLaunches = t->lastLaunches - launches0;
	t->graphRows = t->phiRows;
	t->stepCount = steps0;                                        // nothing has run yet
	t->lastLaunches = launches0;
	t->phiRows = rows0;
	t->rhoParity = parity0;
	t->rhoAll = t->rhoStore + (size_t)parity0 * t->spanDoubles;
	if (rc != PTP_OK || e != cudaSuccess || !graph) {
		if (graph) cudaGraphDestroy(graph);
		cudaGetLastError();
		if (rc == PTP_OK) return ptp_cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
		return rc;
	}
	if (t->graphExec[target]) { cudaGraphExecDestroy(t->graphExec[target]); t->graphExec[target] = nullptr; }
	e = cudaGraphInstantiate(&t->graphExec[target], graph, 0);
	cudaGraphDestroy(graph);
	if (e != cudaSuccess) { t->graphExec[target] = nullptr; return ptp_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
	return PTP_OK;
}

// Replay policy. Forced on / off by ptp_trap_set_graph (or PTP_GRAPH); automatic otherwise: replay where the launches of a
// step are not hidden behind its kernels - loads up to PTP_GRAPH_MAX_RINGS rings on this GPU (default 8 M: a step of a few tens
// of microseconds) and multi-GPU runs (short per-GPU steps plus a barrier every rank waits on).
bool want_graph(ptp_trap* t)
{
	if (t->solver == PTP_SOLVER_SOR || t->plasmas.empty() || t->useGraph == 0) return false;
	if (ptp_comm_size(t) > 1 && !ptp_peer_mode(t)) return false;   // (an NCCL collective inside the step: not captured)
	if (t->useGraph > 0) return true;
	long long rings = 0;
	for (const ptp_plasma* p : t->plasmas) rings += p->nAlive;
	return rings <= t->graphMaxRings || ptp_comm_size(t) > 1;
}

} // namespace

int ptp_trap_step(ptp_trap* t, double dt, int nSteps)
{
	if (!t || nSteps < 0) { ptp_set_error("ptp_trap_step: bad arguments"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	const bool graph = want_graph(t);
	// phase events for every step (up to a bound), so that callers can report the mean kernel time
	const int timed = (!graph && t->phaseEvents && nSteps <= 4096) ? nSteps : 0;
	while ((int)t->evPool.size() < 4 * timed) {
		cudaEvent_t e;
		PTP_CUDA(cudaEventCreate(&e));
		t->evPool.push_back(e);
	}
	t->evSteps = timed;
	PTP_TRY(begin_steps(t));
	PTP_CUDA(cudaEventRecord(t->ev[0], t->stream));
	for (int s = 0; s < nSteps; ++s) {
		if (graph) {
			if (t->graphCfg != t->cfgEpoch || t->graphDt != dt) {       // something a step launches has changed: both parities again
				drop_graph(t);
				t->graphCfg = t->cfgEpoch;
				t->graphDt = dt;
			}
			const int target = t->rhoParity ^ 1;
			if (!t->graphExec[target]) {
				PTP_TRY(capture_step_graph(t, dt, target));
				if (t->graphCfg != t->cfgEpoch) { drop_graph(t); t->graphCfg = t->cfgEpoch; t->graphDt = dt; PTP_TRY(capture_step_graph(t, dt, target)); }   // (the capture's own preparations planned segments)
			}
			if (!t->eNodesValid) PTP_TRY(ptp_node_field(t));      // potentials were replaced since the last step (setPotential, parity hooks)
			PTP_CUDA(cudaGraphLaunch(t->graphExec[target], t->stream));
			t->rhoParity = target;
			t->rhoAll = t->rhoStore + (size_t)target * t->spanDoubles;
			t->lastLaunches += t->graphLaunches;
			t->phiRows = t->graphRows;
			t->eNodesValid = true;
			for (ptp_plasma* p : t->plasmas) p->encValid = true;
			++t->stepCount;
		}
		else PTP_TRY(one_step(t, dt, s < timed ? &t->evPool[4 * s] : nullptr));
		PTP_TRY(maintain_order(t));
	}
	PTP_CUDA(cudaEventRecord(t->ev[4], t->stream));
	return PTP_OK;
}

int ptp_trap_step_programme(ptp_trap* t, double dt, int nSteps, const double* weights)
{
	if (!t || nSteps < 0 || (nSteps > 0 && !weights)) { ptp_set_error("ptp_trap_step_programme: bad arguments"); return PTP_EINVAL; }
	if (t->nBasis < 1) { ptp_set_error("ptp_trap_step_programme: call ptp_trap_set_wall_basis first"); return PTP_ESTATE; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	t->evSteps = 0;
	if (nSteps > 0) PTP_TRY(upload_weights(t, weights, (size_t)nSteps * t->nBasis));
	PTP_TRY(begin_steps(t));
	PTP_CUDA(cudaEventRecord(t->ev[0], t->stream));
	for (int s = 0; s < nSteps; ++s) {
		PTP_TRY(combine_basis(t, t->dWeights + (size_t)s * t->nBasis));   // setPotential(...) of this step; the node field follows in the push
		PTP_TRY(one_step(t, dt, nullptr));
		PTP_TRY(maintain_order(t));
	}
	PTP_CUDA(cudaEventRecord(t->ev[4], t->stream));
	return PTP_OK;
}

int ptp_trap_set_graph(ptp_trap* t, int on)
{
	if (!t) { ptp_set_error("ptp_trap_set_graph: null trap"); return PTP_EINVAL; }
	t->useGraph = on < 0 ? -1 : (on != 0 ? 1 : 0);             // 1: always replay, 0: never, -1: automatic (the default)
	if (!on) drop_graph(t);
	return PTP_OK;
}

int ptp_trap_set_phase_events(ptp_trap* t, int on)
{
	if (!t) { ptp_set_error("ptp_trap_set_phase_events: null trap"); return PTP_EINVAL; }
	t->phaseEvents = on != 0;
	return PTP_OK;
}

int ptp_trap_sync(ptp_trap* t)
{
	if (!t) { ptp_set_error("ptp_trap_sync: null trap"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_trap_last_times(ptp_trap* t, double* ms4)
{
	if (!t || !ms4) { ptp_set_error("ptp_trap_last_times: null argument"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	float whole = 0;
	if (cudaEventElapsedTime(&whole, t->ev[0], t->ev[4]) != cudaSuccess) { cudaGetLastError(); whole = 0; }
	double sum[3] = { 0, 0, 0 };
	for (int s = 0; s < t->evSteps; ++s)
		for (int ph = 0; ph < 3; ++ph) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, t->evPool[4 * s + ph], t->evPool[4 * s + ph + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0; }
			sum[ph] += ms;
		}
	ms4[0] = whole; ms4[1] = sum[0]; ms4[2] = sum[1]; ms4[3] = sum[2];
	if (const char* path = std::getenv("PTP_STEP_TIMES_FILE")) {   // diagnostics: per-step phase times of the last call as CSV
		if (FILE* f = std::fopen(path, "w")) {
			std::fprintf(f, "step,push_deposit_ms,exchange_ms,solve_ms\n");
			for (int s = 0; s < t->evSteps; ++s) {
				float ms[3] = { 0, 0, 0 };
				for (int ph = 0; ph < 3; ++ph)
					if (cudaEventElapsedTime(&ms[ph], t->evPool[4 * s + ph], t->evPool[4 * s + ph + 1]) != cudaSuccess) cudaGetLastError();
				std::fprintf(f, "%d,%.5f,%.5f,%.5f\n", s, ms[0], ms[1], ms[2]);
			}
			std::fclose(f);
		}
	}
	return PTP_OK;
}

int64_t ptp_trap_last_launches(ptp_trap* t) { return t ? t->lastLaunches : 0; }

int ptp_trap_sort(ptp_trap* t)
{
	if (t) ++t->cfgEpoch;
	if (!t) { ptp_set_error("ptp_trap_sort: null trap"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	for (ptp_plasma* p : t->plasmas) PTP_TRY(ptp_sort_plasma(t, p));
	return PTP_OK;
}

int ptp_trap_set_sort_interval(ptp_trap* t, int interval)
{
	if (!t || interval < -1) { ptp_set_error("ptp_trap_set_sort_interval: bad arguments"); return PTP_EINVAL; }
	t->sortInterval = interval;
	t->stepsSinceCheck = 0;
	return PTP_OK;
}

int64_t ptp_trap_sorts_done(ptp_trap* t) { return t ? t->sortsDone : 0; }

int ptp_plasma_set_hot(ptp_plasma* p, int mode)
{
	if (!p || mode < -1 || mode > 1) { ptp_set_error("ptp_plasma_set_hot: bad arguments"); return PTP_EINVAL; }
	if (mode == 1 && !ptp_push_scatter_usable(p->trap)) {
		ptp_set_error("ptp_plasma_set_hot: the per-warp-bin push kernel needs the default tuning (512 threads x 4 rings per thread)");
		return PTP_EINVAL;
	}
	if (mode != p->hot) {
		p->hot = mode;
		p->hotAuto = false;
		p->lastSortStep = -1;
		p->quickSorts = 0;
		p->boundsValid = false;                                  // segment tables are planned per kernel form
		++p->trap->cfgEpoch;
	}
	return PTP_OK;
}

int ptp_plasma_is_hot(ptp_plasma* p) { return p && p->scatter ? 1 : 0; }

int ptp_trap_set_deposit_mode(ptp_trap* t, int mode)
{
	if (t) ++t->cfgEpoch;
	if (!t || (mode != PTP_DEPOSIT_FP64 && mode != PTP_DEPOSIT_FIXED64)) { ptp_set_error("ptp_trap_set_deposit_mode: bad mode"); return PTP_EINVAL; }
	const int old = t->depositMode;
	t->depositMode = mode;
	if (ptp_push_configure(t) != PTP_OK) { t->depositMode = old; return PTP_EINVAL; }
	return PTP_OK;
}

int ptp_trap_set_arith_mode(ptp_trap* t, int mode)
{
	if (t) ++t->cfgEpoch;
	if (!t || (mode != PTP_ARITH_FAST && mode != PTP_ARITH_EXACT)) { ptp_set_error("ptp_trap_set_arith_mode: bad mode"); return PTP_EINVAL; }
	t->arithMode = mode;
	return PTP_OK;
}

int ptp_trap_set_solver(ptp_trap* t, int solver, double sorTolerance, int sorMaxIterations)
{
	if (t) ++t->cfgEpoch;
	if (!t || (solver != PTP_SOLVER_DIRECT && solver != PTP_SOLVER_SOR && solver != PTP_SOLVER_DIRECT_FFT)) { ptp_set_error("ptp_trap_set_solver: bad solver"); return PTP_EINVAL; }
	if (solver == PTP_SOLVER_DIRECT_FFT && !t->fftTw) { ptp_set_error("ptp_trap_set_solver: the FFT inverse needs a power-of-two Nz"); return PTP_EINVAL; }
	t->solver = solver;
	if (sorTolerance > 0) t->sorTol = sorTolerance;
	if (sorMaxIterations > 0) t->sorMaxIter = sorMaxIterations;
	return PTP_OK;
}

int ptp_trap_set_tuning(ptp_trap* t, int threads, int window, int ctas, int ringsPerThread)
{
	if (t) ++t->cfgEpoch;
	if (!t) { ptp_set_error("ptp_trap_set_tuning: null trap"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	const int oT = t->threads, oW = t->window, oC = t->ctas, oR = t->ringsPerThread;
	if (threads > 0) t->threads = threads;
	if (window > 0) t->window = window;
	if (ctas >= 0) t->ctas = ctas;
	if (ringsPerThread > 0) t->ringsPerThread = ringsPerThread;
	if (ptp_push_configure(t) != PTP_OK) { t->threads = oT; t->window = oW; t->ctas = oC; t->ringsPerThread = oR; return PTP_EINVAL; }
	for (ptp_plasma* p : t->plasmas)
		if (p->cap) PTP_TRY(ptp_build_segments(t, p));
	return PTP_OK;
}

int ptp_plasma_create(ptp_trap* t, ptp_plasma** out, double mass, double charge)
{
	if (t) ++t->cfgEpoch;
	if (!t || !out || !(mass > 0) || charge == 0) { ptp_set_error("ptp_plasma_create: bad arguments"); return PTP_EINVAL; }
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	PTP_TRY(ensure_species_capacity(t, (int)t->plasmas.size() + 1));
	ptp_plasma* p = new ptp_plasma;
	p->trap = t;
	p->index = (int)t->plasmas.size();
	p->mass = mass;
	p->charge = charge;
	p->rowOff.assign(t->Nr + 1, 0);
	p->rowLive.assign(t->Nr, 0);
	p->hot = t->scatterPolicy < -1 || t->scatterPolicy > 1 ? -1 : t->scatterPolicy;
	PTP_CUDA(cudaMalloc(&p->dLost, 2 * sizeof(unsigned long long)));
	PTP_CUDA(cudaMemset(p->dLost, 0, 2 * sizeof(unsigned long long)));
	const size_t gb = (size_t)t->G * sizeof(double);
	PTP_CUDA(cudaMemset(t->rhoAll + (size_t)p->index * t->G, 0, gb));
	PTP_CUDA(cudaMemset(t->phiSelfAll + (size_t)p->index * t->G, 0, gb));
	t->plasmas.push_back(p);
	t->eNodesValid = false;
	*out = p;
	return PTP_OK;
}

int ptp_plasma_destroy(ptp_plasma* p)
{
	if (!p) return PTP_OK;
	ptp_trap* t = p->trap;
	++t->cfgEpoch; ++t->layoutEpoch;
	cudaSetDevice(t->device);
	cudaStreamSynchronize(t->stream);
	// later species move down one slice so that slices stay contiguous and in registration order
	const size_t gb = (size_t)t->G * sizeof(double);
	for (size_t s = p->index + 1; s < t->plasmas.size(); ++s) {
		cudaMemcpy(t->rhoAll + (s - 1) * t->G, t->rhoAll + s * t->G, gb, cudaMemcpyDeviceToDevice);
		cudaMemcpy(t->phiSelfAll + (s - 1) * t->G, t->phiSelfAll + s * t->G, gb, cudaMemcpyDeviceToDevice);
		cudaMemcpy(t->dScale + (s - 1), t->dScale + s, sizeof(double), cudaMemcpyDeviceToDevice);
		t->plasmas[s]->index = (int)s - 1;
	}
	t->plasmas.erase(t->plasmas.begin() + p->index);
	for (ptp_plasma* q : t->plasmas) q->encValid = false;       // the touched-node ranges were not moved with the grids
	t->eNodesValid = false;
	cudaFree(p->z); cudaFree(p->v); cudaFree(p->id); cudaFree(p->zAlt); cudaFree(p->vAlt); cudaFree(p->idAlt); cudaFree(p->sortScratch); cudaFree(p->planScratch);
	cudaFree(p->dRowOff); cudaFree(p->dSegs); cudaFree(p->dCtaSegBegin); cudaFree(p->dSegBounds); cudaFree(p->dLost);
	cudaFree(p->dLossLog); cudaFree(p->vSaved); cudaFree(p->vSavedAlt);
	delete p;
	return PTP_OK;
}

int ptp_plasma_deposit(ptp_plasma* p)
{
	if (!p) { ptp_set_error("ptp_plasma_deposit: null plasma"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	t->lastLaunches = 0;
	PTP_TRY(ptp_layout_sync(t));                                 // row extent and fixed-point scale (collective after a (re)load)
	if (!p->boundsValid) PTP_TRY(ptp_bounds_launch(t, p));
	void* rho = t->rhoAll + (size_t)p->index * t->G;
	PTP_CUDA(cudaMemsetAsync(rho, 0, (size_t)t->G * sizeof(double), t->stream));
	PTP_CUDA(cudaMemsetAsync(t->rhoAll + (size_t)t->capS * t->G + (size_t)p->index * t->Nr, 0, (size_t)t->Nr * sizeof(double), t->stream));
	PTP_TRY(ptp_push_launch(t, p, 0.0, false));
	// the kernel saw this rank's rings only: its touched-node ranges describe the grid on one GPU, not the all-reduced one
	p->encValid = ptp_comm_size(t) == 1;
	return ptp_comm_allreduce(t, rho, (size_t)t->rowExtent * (t->Nz + 1), t->depositMode == PTP_DEPOSIT_FIXED64);
}

int ptp_plasma_deposit_solve(ptp_plasma* p)
{
	PTP_TRY(ptp_plasma_deposit(p));
	ptp_trap* t = p->trap;
	PTP_TRY(solve_species(t, p->index, 1));
	t->eNodesValid = false;
	return PTP_OK;
}

int ptp_plasma_get_rhs(ptp_plasma* p, double* rhs)
{
	if (!p || !rhs) { ptp_set_error("ptp_plasma_get_rhs: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	const size_t gb = (size_t)t->G * sizeof(double);
	PTP_CUDA(cudaMemcpyAsync(rhs, t->rhoAll + (size_t)p->index * t->G, gb, cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	const double scale = -p->macroChargeDensity / 8.8541878128e-12;
	if (t->depositMode == PTP_DEPOSIT_FIXED64) {
		const double inv = 1.0 / (double)(1ULL << t->fixedBits);
		for (long long i = 0; i < t->G; ++i) {
			long long w;
			std::memcpy(&w, &rhs[i], sizeof(w));
			rhs[i] = ((double)w * inv) * scale;
		}
	}
	else for (long long i = 0; i < t->G; ++i) rhs[i] *= scale;
	return PTP_OK;
}

int ptp_plasma_get_self_potential(ptp_plasma* p, double* phi)
{
	if (!p || !phi) { ptp_set_error("ptp_plasma_get_self_potential: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_TRY(ptp_materialize_fields(t));
	PTP_CUDA(cudaMemcpyAsync(phi, t->phiSelfAll + (size_t)p->index * t->G, (size_t)t->G * sizeof(double), cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	return PTP_OK;
}

int ptp_plasma_set_self_potential(ptp_plasma* p, const double* phi)
{
	if (!p || !phi) { ptp_set_error("ptp_plasma_set_self_potential: null argument"); return PTP_EINVAL; }
	ptp_trap* t = p->trap;
	PTP_CUDA(cudaSetDevice(t->device));
	PTP_TRY(ptp_materialize_fields(t));                          // the other species' grids in full before this one is replaced
	PTP_CUDA(cudaMemcpyAsync(t->phiSelfAll + (size_t)p->index * t->G, phi, (size_t)t->G * sizeof(double), cudaMemcpyHostToDevice, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	t->eNodesValid = false;
	return PTP_OK;
}

} // extern "C"
