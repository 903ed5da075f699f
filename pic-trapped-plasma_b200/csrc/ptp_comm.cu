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

struct PtpComm {
	ncclComm_t comm = nullptr;
	int nRanks = 1, rank = 0;
};

namespace {
struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
	g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.GetErrorString;
	if (!g_nccl.ok) ptp_set_error("NCCL library lacks required symbols");
	return g_nccl.ok;
}

int nccl_fail(ncclResult_t r, const char* what)
{
	ptp_set_error(std::string(what) + ": " + g_nccl.GetErrorString(r));
	return PTP_ECOMM;
}
} // namespace

int ptp_comm_size(ptp_trap* t) { return t->comm ? t->comm->nRanks : 1; }

int ptp_comm_allreduce(ptp_trap* t, void* buf, size_t count, bool isInt64)
{
	if (!t->comm || t->comm->nRanks == 1) return PTP_OK;
	ncclResult_t r = g_nccl.AllReduce(buf, buf, count, isInt64 ? ncclInt64 : ncclFloat64, ncclSum, t->comm->comm, t->stream);
	if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
	return PTP_OK;
}

void ptp_comm_free(ptp_trap* t)
{
	if (t->comm) {
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
	if (!t || kind < 0 || kind > 1) { ptp_set_error("ptp_trap_set_allreduce: bad arguments"); return PTP_EINVAL; }
	if (kind == 1) { ptp_set_error("peer-memory all-reduce not available in this build"); return PTP_EINVAL; }
	t->allreduceKind = kind;
	return PTP_OK;
}

} // extern "C"
