// Multi-GPU exchange step: one process per GPU, rings sharded, rank-local deposit grids summed.
//
// The reference is single-threaded and has no communication at all; this is the one exchange step the
// sharded step needs: all-reduce(sum) of the species' deposit grids after K1, fp64 or int64
// (fixed-point mode: the sum is then bitwise independent of the rank count). NCCL is resolved at run time
// with dlopen so that (i) the library has no link-time NCCL dependency for single-GPU use and (ii) inside a
// torch process the already-loaded torch-bundled libnccl.so.2 is the one that is used.
#include "ptp_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <limits.h>

#include <algorithm>
#include <cstring>

struct PtpComm {
	ncclComm_t comm = nullptr;
	int nRanks = 1, rank = 0;
	// peer-memory mode: every rank's rhoStore mapped into this process (own entry = own pointer)
	double* peerBase[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	bool mapped = false;
	size_t spanDoubles = 0;          // capS * G at mapping time
	unsigned long long* dEpoch = nullptr; // barrier generation, kept on the device so that barrier launches can be replayed from a CUDA graph;
	                                      // all ranks advance it in lock-step
	// gather exchange (ptp_trap_set_allreduce(t, 3)): one allocation per rank, mapped by every rank:
	//   [2 parities][nRanks][span] deposit grids as pushed by each rank | [nRanks][PTP_EXCHANGE_CTAS] flags | [2] exchange count, CTA ticket
	double* gatherStore = nullptr;
	size_t gatherSpan = 0;
	double* peerGather[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	bool gatherMapped = false;
};

namespace {
struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
} g_nccl;

bool load_nccl()
{
	if (g_nccl.ok) return true;
	const char* names[] = { "libnccl.so.2", "libnccl.so" };
	for (const char* n : names) {
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) break;
	}
	if (!g_nccl.handle) { ptp_set_error(std::string("cannot load NCCL: ") + dlerror()); return false; }
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
	g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
	g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
	g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather && g_nccl.GetErrorString;
	if (!g_nccl.ok) ptp_set_error("NCCL library lacks required symbols");
	return g_nccl.ok;
}

int nccl_fail(ncclResult_t r, const char* what)
{
	ptp_set_error(std::string(what) + ": " + g_nccl.GetErrorString(r));
	return PTP_ECOMM;
}

// Flag barrier over peer memory: rank r stores the epoch into slot r of every rank's flag array (release, system scope),
// then spins on its own array until every slot has reached the epoch (acquire). Launched after the push kernel in stream
// order, so this rank's remote adds are complete before its flag can be seen.
struct PeerFlags { unsigned long long* f[8]; };
__global__ void k_peer_barrier(PeerFlags flags, int rank, int nRanks, unsigned long long* dEpoch)
{
	__shared__ unsigned long long sEpoch;
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();                                             // this rank's push kernels (and their remote adds) are complete
	if (threadIdx.x == 0) { sEpoch = *dEpoch + 1; *dEpoch = sEpoch; }
	__syncthreads();
	const unsigned long long epoch = sEpoch;
	const int p = threadIdx.x;
	if (p < nRanks) {
		__threadfence_system();
		unsigned long long* remote = flags.f[p] + rank;
		asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(remote), "l"(epoch) : "memory");
		const unsigned long long* mine = flags.f[rank] + p;
		unsigned long long seen;
		do {
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(seen) : "l"(mine) : "memory");
		} while (seen < epoch);
	}
	__syncthreads();
	__threadfence_system();
}

// ---- gather exchange ---------------------------------------------------------------------------------------------
// Every rank pushes its step's deposit grids (the populated rows and the touched-node ranges, local sums of its own rings)
// into slot [rank] of EVERY rank's gather area with plain stores over NVLink, signals, waits for the other ranks' signals
// and then sums the slots in rank order into its own grids: an all-gather + local reduction. Compared with adding into the
// peers' grids from the push kernel's flush: no remote atomics at all (7 x 56 KB of posted stores per rank and step on the
// default grid instead of ~70 000 contended remote adds), the push kernel is the single-GPU kernel, and the sum has a fixed
// order - every rank holds bitwise the same grids in fp64 mode too. The work is cut into PTP_EXCHANGE_CTAS independent slices
// (one CTA each, its own flags), so there is no grid-wide synchronisation inside the kernel.
constexpr int PTP_EXCHANGE_CTAS = 8;
constexpr int PTP_EXCHANGE_THREADS = 512;

struct ExchangeArgs {
	double* L;                       // this rank's deposit grids of the step: local sums on entry, global sums on exit
	double* G[8];                    // gather area of every rank for this step's parity: [nRanks][span]
	unsigned long long* flags[8];    // flag array of every rank: [nRanks][PTP_EXCHANGE_CTAS]
	unsigned long long* epoch;       // own: [0] exchanges completed, [1] CTA ticket
	long long span, gridDoubles, rowWords;   // doubles per slot; doubles per species grid; populated rows x (Nz + 1)
	int rank, nRanks, nS, capS, Nr, fixed, rows, pad;   // rows = populated rows
};

// One warp per (species, row) of the populated rows; a row travels as its touched node range only (the ranges the push
// kernels publish: a few hundred of the 4097 nodes of a fine-grid row), together with the range itself.
__global__ void __launch_bounds__(PTP_EXCHANGE_THREADS) k_peer_exchange(const ExchangeArgs a)
{
	ptp_pdl_launch_dependents();
	ptp_pdl_wait();                                             // this rank's push kernels are complete
	const int tid = threadIdx.x, lane = tid & 31, b = blockIdx.x;
	const int warpsPerCta = PTP_EXCHANGE_THREADS / 32;
	const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(a.epoch) + 1;
	const int n1 = (int)(a.rowWords / max(a.rows, 1)), rowsTotal = a.nS * a.rows;
	auto decode = [n1](unsigned long long w, int& lo, int& hi) {    // encoded maxima (Nz + 2 - kmin, kmax + 1 + 1); 0 = untouched
		const unsigned int x = (unsigned int)(w & 0xffffffffULL), y = (unsigned int)(w >> 32);
		if (y == 0) { lo = 1; hi = 0; }
		else { lo = n1 + 1 - (int)x; hi = (int)y - 1; }
	};
	unsigned long long* L64 = reinterpret_cast<unsigned long long*>(a.L);
	// push this rank's rows (touched range + the range) into slot [rank] of every rank's gather area
	for (int w = b * warpsPerCta + (tid >> 5); w < rowsTotal; w += gridDim.x * warpsPerCta) {
		const int s = w / a.rows, j = w - s * a.rows;
		const long long bOff = (long long)a.capS * a.gridDoubles + (long long)s * a.Nr + j, gOff = (long long)s * a.gridDoubles + (long long)j * n1;
		const unsigned long long bnd = L64[bOff];
		int lo, hi;
		decode(bnd, lo, hi);
		if (lane == 0)
			for (int p = 0; p < a.nRanks; ++p) reinterpret_cast<unsigned long long*>(a.G[p] + (long long)a.rank * a.span)[bOff] = bnd;
		for (int k = lo + lane; k <= hi; k += 32) {
			const unsigned long long val = L64[gOff + k];           // read once, stored to every rank's slot (posted writes)
#pragma unroll
			for (int p = 0; p < 8; ++p)
				if (p < a.nRanks) reinterpret_cast<unsigned long long*>(a.G[p] + (long long)a.rank * a.span)[gOff + k] = val;
		}
	}
	__syncthreads();
	if (tid < a.nRanks) {
		__threadfence_system();                                 // the CTA's stores (ordered by the barrier) before the flag
		unsigned long long* remote = a.flags[tid] + a.rank * PTP_EXCHANGE_CTAS + b;
		asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(remote), "l"(epoch) : "memory");
		const unsigned long long* mine = a.flags[a.rank] + tid * PTP_EXCHANGE_CTAS + b;
		unsigned long long seen;
		do {
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(seen) : "l"(mine) : "memory");
		} while (seen < epoch);
	}
	__syncthreads();
	// sum the slots in rank order (L2 loads: the lines were written by the peers) into this rank's grids
	const unsigned long long* mineG = reinterpret_cast<const unsigned long long*>(a.G[a.rank]);
	for (int w = b * warpsPerCta + (tid >> 5); w < rowsTotal; w += gridDim.x * warpsPerCta) {
		const int s = w / a.rows, j = w - s * a.rows;
		const long long bOff = (long long)a.capS * a.gridDoubles + (long long)s * a.Nr + j, gOff = (long long)s * a.gridDoubles + (long long)j * n1;
		// every load of a round is issued before the first value is used (a load under a data-dependent branch followed by a
		// dependent add would serialise eight L2 round trips per node: measured 33 us per exchange at 8 ranks, 77 us on the fine grid)
		unsigned long long bw[8];
#pragma unroll
		for (int r = 0; r < 8; ++r) bw[r] = r < a.nRanks ? __ldcg(mineG + (long long)r * a.span + bOff) : 0ULL;
		int lo[8], hi[8], ulo = INT_MAX, uhi = INT_MIN;
		unsigned int mx = 0, my = 0;
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			decode(bw[r], lo[r], hi[r]);
			mx = max(mx, (unsigned int)(bw[r] & 0xffffffffULL));
			my = max(my, (unsigned int)(bw[r] >> 32));
			if (lo[r] <= hi[r]) { ulo = min(ulo, lo[r]); uhi = max(uhi, hi[r]); }
		}
		if (lane == 0) L64[bOff] = ((unsigned long long)my << 32) | mx;
		for (int k0 = ulo + lane; k0 <= uhi; k0 += 64) {           // two nodes per lane and round: 16 loads in flight
			unsigned long long w8[2][8];
#pragma unroll
			for (int u = 0; u < 2; ++u) {
				const int k = k0 + 32 * u;
#pragma unroll
				for (int r = 0; r < 8; ++r) w8[u][r] = (k <= uhi && k >= lo[r] && k <= hi[r]) ? __ldcg(mineG + (long long)r * a.span + gOff + k) : 0ULL;
			}
#pragma unroll
			for (int u = 0; u < 2; ++u) {
				const int k = k0 + 32 * u;
				if (k > uhi) continue;
				unsigned long long out;
				if (a.fixed) {
					out = 0ULL;
#pragma unroll
					for (int r = 0; r < 8; ++r) out += w8[u][r];
				}
				else {
					double sum = 0.0;                               // rank order; a rank whose range does not cover the node adds +0.0
#pragma unroll
					for (int r = 0; r < 8; ++r) sum = __dadd_rn(sum, __longlong_as_double((long long)w8[u][r]));
					out = (unsigned long long)__double_as_longlong(sum);
				}
				L64[gOff + k] = out;
			}
		}
	}
	__syncthreads();
	if (tid == 0) {                                             // every CTA has read the exchange count before the last one to finish advances it
		__threadfence();
		const unsigned long long ticket = atomicAdd(a.epoch + 1, 1ULL);
		if (ticket == (unsigned long long)gridDim.x - 1) { a.epoch[1] = 0ULL; a.epoch[0] = epoch; }
	}
}

void unmap_gather(ptp_trap* t)
{
	PtpComm* c = t->comm;
	if (!c || !c->gatherMapped) return;
	for (int r = 0; r < c->nRanks; ++r)
		if (r != c->rank && c->peerGather[r]) cudaIpcCloseMemHandle(c->peerGather[r]);
	for (auto& b : c->peerGather) b = nullptr;
	c->gatherMapped = false;
}

void unmap_peers(ptp_trap* t)
{
	PtpComm* c = t->comm;
	if (!c || !c->mapped) return;
	for (int r = 0; r < c->nRanks; ++r)
		if (r != c->rank && c->peerBase[r]) cudaIpcCloseMemHandle(c->peerBase[r]);
	for (auto& b : c->peerBase) b = nullptr;
	c->mapped = false;
}
} // namespace

int ptp_comm_size(ptp_trap* t) { return t->comm ? t->comm->nRanks : 1; }
int ptp_comm_rank(ptp_trap* t) { return t->comm ? t->comm->rank : 0; }

bool ptp_peer_fused(ptp_trap* t) { return t->comm && t->comm->nRanks > 1 && t->allreduceKind == 1; }
bool ptp_peer_gather(ptp_trap* t) { return t->comm && t->comm->nRanks > 1 && t->allreduceKind == 3; }
bool ptp_peer_mode(ptp_trap* t) { return ptp_peer_fused(t) || ptp_peer_gather(t); }

namespace {
// Collective: (re)allocate this rank's gather area, exchange the IPC handles, map the peers' areas.
int gather_prepare(ptp_trap* t)
{
	PtpComm* c = t->comm;
	if (!t->peerStale && c->gatherMapped && c->gatherSpan == t->spanDoubles) return PTP_OK;
	if (c->nRanks > 8) { ptp_set_error("peer-memory mode supports at most 8 ranks"); return PTP_EINVAL; }
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	unmap_gather(t);
	cudaFree(c->gatherStore);
	c->gatherStore = nullptr;
	const size_t span = t->spanDoubles;
	const size_t tailWords = (size_t)c->nRanks * PTP_EXCHANGE_CTAS + 2;
	const size_t bytes = (2 * (size_t)c->nRanks * span + tailWords) * sizeof(double);
	PTP_CUDA(cudaMalloc(&c->gatherStore, bytes));
	PTP_CUDA(cudaMemsetAsync(c->gatherStore, 0, bytes, t->stream));
	cudaIpcMemHandle_t mine;
	PTP_CUDA(cudaIpcGetMemHandle(&mine, c->gatherStore));
	unsigned char* dH = nullptr;
	const size_t hs = sizeof(cudaIpcMemHandle_t);
	PTP_CUDA(cudaMalloc(&dH, hs * c->nRanks));
	PTP_CUDA(cudaMemcpyAsync(dH + hs * c->rank, &mine, hs, cudaMemcpyHostToDevice, t->stream));
	ncclResult_t r = g_nccl.AllGather(dH + hs * c->rank, dH, hs, ncclChar, c->comm, t->stream);   // (behind the zeroing: nobody signals into a dirty flag array)
	if (r != ncclSuccess) { cudaFree(dH); return nccl_fail(r, "ncclAllGather(ipc handles)"); }
	std::vector<cudaIpcMemHandle_t> all(c->nRanks);
	PTP_CUDA(cudaMemcpyAsync(all.data(), dH, hs * c->nRanks, cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	cudaFree(dH);
	for (int p = 0; p < c->nRanks; ++p) {
		if (p == c->rank) { c->peerGather[p] = c->gatherStore; continue; }
		void* ptr = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) return ptp_cuda_fail(e, "cudaIpcOpenMemHandle (peer-memory mode needs P2P-capable GPUs)", __FILE__, __LINE__);
		c->peerGather[p] = static_cast<double*>(ptr);
	}
	c->gatherMapped = true;
	c->gatherSpan = span;
	t->peerStale = false;
	t->peerCleanEpoch = -1;
	return PTP_OK;
}
} // namespace

int ptp_peer_exchange(ptp_trap* t)
{
	PtpComm* c = t->comm;
	ExchangeArgs a{};
	const size_t span = c->gatherSpan;
	a.L = t->rhoAll;
	for (int p = 0; p < c->nRanks; ++p) {
		a.G[p] = c->peerGather[p] + (size_t)t->rhoParity * c->nRanks * span;
		a.flags[p] = reinterpret_cast<unsigned long long*>(c->peerGather[p] + 2 * (size_t)c->nRanks * span);
	}
	a.epoch = reinterpret_cast<unsigned long long*>(c->gatherStore + 2 * (size_t)c->nRanks * span) + (size_t)c->nRanks * PTP_EXCHANGE_CTAS;
	a.span = (long long)span;
	a.gridDoubles = t->G;
	a.rowWords = (long long)std::min(t->rowExtent, t->Nr) * (t->Nz + 1);
	a.rank = c->rank; a.nRanks = c->nRanks; a.nS = (int)t->plasmas.size(); a.capS = t->capS; a.Nr = t->Nr;
	a.rows = std::min(t->rowExtent, t->Nr);
	a.fixed = t->depositMode == PTP_DEPOSIT_FIXED64 ? 1 : 0;
	cudaError_t e = ptp_launch(k_peer_exchange, dim3(PTP_EXCHANGE_CTAS), dim3(PTP_EXCHANGE_THREADS), 0, t->stream, t->usePdl, a);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_peer_exchange launch", __FILE__, __LINE__);
	t->lastLaunches++;
	return PTP_OK;
}

// Collective over all ranks: exchange the CUDA IPC handles of the rhoStore allocations (through an NCCL all-gather)
// and map every peer's allocation. Needed once, and again whenever a rank had to reallocate (more species).
int ptp_peer_prepare(ptp_trap* t)
{
	PtpComm* c = t->comm;
	if (t->allreduceKind == 3) return gather_prepare(t);
	// Remapping is a collective (all-gather of the IPC handles): ranks must agree on when it happens. They do when every rank
	// issues the same sequence of calls (SPMD use: same species created in the same order), which include/ptp.h requires of
	// multi-rank callers - the allocation is replaced only when a species is added beyond the reserved capacity.
	if (!t->peerStale && c->mapped) return PTP_OK;
	int* dFlag = nullptr;
	ncclResult_t r;
	if (c->nRanks > 8) { ptp_set_error("peer-memory mode supports at most 8 ranks"); return PTP_EINVAL; }
	unmap_peers(t);
	cudaIpcMemHandle_t mine;
	PTP_CUDA(cudaIpcGetMemHandle(&mine, t->rhoStore));
	unsigned char* dH = nullptr;
	const size_t hs = sizeof(cudaIpcMemHandle_t);
	PTP_CUDA(cudaMalloc(&dH, hs * c->nRanks));
	PTP_CUDA(cudaMemcpyAsync(dH + hs * c->rank, &mine, hs, cudaMemcpyHostToDevice, t->stream));
	r = g_nccl.AllGather(dH + hs * c->rank, dH, hs, ncclChar, c->comm, t->stream);
	if (r != ncclSuccess) { cudaFree(dH); return nccl_fail(r, "ncclAllGather(ipc handles)"); }
	std::vector<cudaIpcMemHandle_t> all(c->nRanks);
	PTP_CUDA(cudaMemcpyAsync(all.data(), dH, hs * c->nRanks, cudaMemcpyDeviceToHost, t->stream));
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	cudaFree(dH);
	for (int p = 0; p < c->nRanks; ++p) {
		if (p == c->rank) { c->peerBase[p] = t->rhoStore; continue; }
		void* ptr = nullptr;
		cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) return ptp_cuda_fail(e, "cudaIpcOpenMemHandle (peer-memory mode needs P2P-capable GPUs)", __FILE__, __LINE__);
		c->peerBase[p] = static_cast<double*>(ptr);
	}
	c->mapped = true;
	c->spanDoubles = t->spanDoubles;
	t->peerStale = false;
	t->peerCleanEpoch = -1;
	// every rank restarts its barrier generation and flags at zero; nobody may signal before everybody has done so
	if (!c->dEpoch) PTP_CUDA(cudaMalloc(&c->dEpoch, sizeof(unsigned long long)));
	PTP_CUDA(cudaMemsetAsync(c->dEpoch, 0, sizeof(unsigned long long), t->stream));
	PTP_CUDA(cudaMemsetAsync(t->rhoStore + 2 * t->spanDoubles, 0, 64 * sizeof(unsigned long long), t->stream));
	PTP_CUDA(cudaMalloc(&dFlag, sizeof(int)));
	PTP_CUDA(cudaMemsetAsync(dFlag, 0, sizeof(int), t->stream));
	r = g_nccl.AllReduce(dFlag, dFlag, 1, ncclInt32, ncclMax, c->comm, t->stream);
	if (r != ncclSuccess) { cudaFree(dFlag); return nccl_fail(r, "ncclAllReduce(barrier reset)"); }
	PTP_CUDA(cudaStreamSynchronize(t->stream));
	cudaFree(dFlag);
	return PTP_OK;
}

void ptp_peer_targets(ptp_trap* t, int parity, size_t offsetDoubles, void** out, int* n)
{
	PtpComm* c = t->comm;
	*n = c->nRanks;
	for (int p = 0; p < c->nRanks; ++p) out[p] = c->peerBase[p] + (size_t)parity * c->spanDoubles + offsetDoubles;
}

int ptp_peer_barrier(ptp_trap* t)
{
	PtpComm* c = t->comm;
	PeerFlags f{};
	for (int p = 0; p < c->nRanks; ++p) f.f[p] = reinterpret_cast<unsigned long long*>(c->peerBase[p] + 2 * c->spanDoubles);
	cudaError_t e = ptp_launch(k_peer_barrier, dim3(1), dim3(32), 0, t->stream, t->usePdl, f, c->rank, c->nRanks, c->dEpoch);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "k_peer_barrier launch", __FILE__, __LINE__);
	t->lastLaunches++;
	return PTP_OK;
}

int ptp_comm_allreduce(ptp_trap* t, void* buf, size_t count, bool isInt64)
{
	if (!t->comm || t->comm->nRanks == 1) return PTP_OK;
	ncclResult_t r = g_nccl.AllReduce(buf, buf, count, isInt64 ? ncclInt64 : ncclFloat64, ncclSum, t->comm->comm, t->stream);
	if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
	return PTP_OK;
}

static int comm_reduce_int(ptp_trap* t, int* value, int n, ncclRedOp_t op)
{
	if (!t->comm || t->comm->nRanks == 1) return PTP_OK;
	int* dV = nullptr;
	PTP_CUDA(cudaMalloc(&dV, n * sizeof(int)));
	cudaError_t e = cudaMemcpyAsync(dV, value, n * sizeof(int), cudaMemcpyHostToDevice, t->stream);
	if (e != cudaSuccess) { cudaFree(dV); return ptp_cuda_fail(e, "cudaMemcpyAsync", __FILE__, __LINE__); }
	ncclResult_t r = g_nccl.AllReduce(dV, dV, n, ncclInt32, op, t->comm->comm, t->stream);
	if (r != ncclSuccess) { cudaFree(dV); return nccl_fail(r, "ncclAllReduce(int)"); }
	e = cudaMemcpyAsync(value, dV, n * sizeof(int), cudaMemcpyDeviceToHost, t->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
	cudaFree(dV);
	if (e != cudaSuccess) return ptp_cuda_fail(e, "ptp_comm_max_int", __FILE__, __LINE__);
	return PTP_OK;
}

int ptp_comm_max_int(ptp_trap* t, int* value, int n) { return comm_reduce_int(t, value, n, ncclMax); }
int ptp_comm_sum_int(ptp_trap* t, int* value, int n) { return comm_reduce_int(t, value, n, ncclSum); }

// Quantities every rank must agree on after a (re)load, settled by one small collective per load: the outermost populated
// row (rows above it never receive a deposit) and the fixed-point scale 2^F of the deposit sums. F follows from the GLOBAL
// ring count (F = min(40, 62 - ceil(log2(N + 1))): node sums stay below 2^62); shards are unequal, so every rank gets the
// ring count as the sum over the ranks - the same F as a single GPU holding the whole load would choose, so that results stay
// bitwise independent of the rank count (a per-rank estimate would differ between ranks near a power of two).
int ptp_layout_sync(ptp_trap* t)
{
	if (t->extentEpoch == t->layoutEpoch) return PTP_OK;
	int ext = 0;
	long long total = 0;
	for (const ptp_plasma* p : t->plasmas) {
		total += p->nUploaded;
		for (int j = (int)p->rowLive.size() - 1; j >= ext; --j)
			if (p->rowLive[j] > 0) { ext = j + 1; break; }
	}
	if (t->comm && t->comm->nRanks > 1) {
		PTP_TRY(ptp_comm_max_int(t, &ext));
		// the global ring count as the sum of the shards (three 20-bit digits through the int32 collective)
		int digit[3] = { (int)(total & 0xfffff), (int)((total >> 20) & 0xfffff), (int)(total >> 40) };
		PTP_TRY(ptp_comm_sum_int(t, digit, 3));
		total = ((long long)digit[2] << 40) + ((long long)digit[1] << 20) + (long long)digit[0];
	}
	int bits = 0;
	while ((1LL << bits) < total + 1) ++bits;
	const int fixedBits = std::min(40, 62 - bits);
	if (fixedBits != t->fixedBits) { t->fixedBits = fixedBits; ++t->cfgEpoch; }
	t->rowExtent = ext;
	t->extentEpoch = t->layoutEpoch;
	return PTP_OK;
}

int ptp_row_extent(ptp_trap* t, int* extent)
{
	PTP_TRY(ptp_layout_sync(t));
	*extent = t->rowExtent;
	return PTP_OK;
}

void ptp_comm_free(ptp_trap* t)
{
	if (t->comm) {
		unmap_peers(t);
		unmap_gather(t);
		cudaFree(t->comm->gatherStore);
		cudaFree(t->comm->dEpoch);
		if (t->comm->comm && g_nccl.ok) g_nccl.CommDestroy(t->comm->comm);
		delete t->comm;
		t->comm = nullptr;
	}
}

extern "C" {

int ptp_comm_unique_id(void* id128)
{
	if (!id128) { ptp_set_error("ptp_comm_unique_id: null buffer"); return PTP_EINVAL; }
	if (!load_nccl()) return PTP_ECOMM;
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
	memcpy(id128, &id, sizeof(id));
	return PTP_OK;
}

int ptp_trap_comm_init(ptp_trap* t, const void* id128, int nRanks, int rank)
{
	if (!t || !id128 || nRanks < 1 || rank < 0 || rank >= nRanks) { ptp_set_error("ptp_trap_comm_init: bad arguments"); return PTP_EINVAL; }
	if (!load_nccl()) return PTP_ECOMM;
	PTP_CUDA(cudaSetDevice(t->device));
	++t->cfgEpoch;
	ptp_comm_free(t);
	t->comm = new PtpComm;
	t->comm->nRanks = nRanks;
	t->comm->rank = rank;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclResult_t r = g_nccl.CommInitRank(&t->comm->comm, nRanks, id, rank);
	if (r != ncclSuccess) { delete t->comm; t->comm = nullptr; return nccl_fail(r, "ncclCommInitRank"); }
	return PTP_OK;
}

int ptp_trap_set_allreduce(ptp_trap* t, int kind)
{
	if (!t || kind < 0 || kind > 3) { ptp_set_error("ptp_trap_set_allreduce: bad arguments"); return PTP_EINVAL; }
	if (kind >= 1 && !t->comm) { ptp_set_error("ptp_trap_set_allreduce: peer-memory mode needs ptp_trap_comm_init first"); return PTP_ESTATE; }
	// auto, by measurement (8 x B200, profiles/r02_scale.txt): on the default grid the fused adds cost the push kernel ~7 us but the
	// exchange is then one flag barrier (15 us with the skew between the ranks) against 23 us for the gather kernel - 0.109 vs
	// 0.113 ms per step of c4 at 8 GPUs, equal at 2; on the 4096 x 1024 grid, whose sparse plasma tails deposit outside the
	// private windows, every such ring would cost two remote atomics per peer (round 1: 1.3 ms per push) - gather there.
	if (kind == 2) kind = t->G <= (1LL << 20) ? 1 : 3;
	if (kind != t->allreduceKind) { t->peerStale = true; t->peerCleanEpoch = -1; }
	t->allreduceKind = kind;
	++t->cfgEpoch;
	return PTP_OK;
}

} // extern "C"
