// peer.cu -- device buffers that the other ranks of a single-node job can read directly over NVLink / NVSwitch
// (CUDA IPC: cudaMalloc + cudaIpcGetMemHandle on the owner, cudaIpcOpenMemHandle on the peers).
//
// Used for the loser store of the look-ahead cache (model_no_ddp.py:176-179: the rows of ids that miss in the
// forward).  The un-cacheable ids of a window are the same on every rank (every rank runs the same deterministic
// plan on the global window), so the store is SHARDED: rank r prefetches 1/W of the rows from the host master over
// its own PCIe link and every rank's forward reads a missing row from the HBM of the rank that holds it.  Compared
// with one full store per rank this cuts the PCIe prefetch and the HBM footprint W times (42 GB -> 5 GB per store at
// 8 GPUs), and no miss has to fall back to a dependent zero-copy PCIe read inside the forward any more.
#include "common.cuh"

extern "C" int cdlrm_peer_alloc(int device, int64_t bytes, void** d_ptr, void* handle_out) {
    ARG_CHECK(bytes > 0 && d_ptr && handle_out);
    CU_CHECK(cudaSetDevice(device));
    void* p = nullptr;
    CU_CHECK(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        cdlrm_set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return CDLRM_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == CDLRM_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    *d_ptr = p;
    return CDLRM_OK;
}

extern "C" int cdlrm_peer_open(int device, const void* handle, void** d_ptr) {
    ARG_CHECK(handle && d_ptr);
    CU_CHECK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CU_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CDLRM_OK;
}

extern "C" int cdlrm_peer_close(int device, void* d_ptr) {
    if (!d_ptr) return CDLRM_OK;
    CU_CHECK(cudaSetDevice(device));
    CU_CHECK(cudaIpcCloseMemHandle(d_ptr));
    return CDLRM_OK;
}

extern "C" int cdlrm_peer_free(int device, void* d_ptr) {
    if (!d_ptr) return CDLRM_OK;
    CU_CHECK(cudaSetDevice(device));
    CU_CHECK(cudaFree(d_ptr));
    return CDLRM_OK;
}
