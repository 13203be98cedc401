// CUDA-IPC peer buffers (one process per GPU, single node, NVLink/NVSwitch).
//
// The reference moves K/V and dK/dV with NCCL isend/irecv, all_gather and reduce_scatter
// (/root/reference/ring_flash_attn/utils.py:98-168).  Here communication is plain stores/loads on
// peer-mapped memory issued by the attention kernels themselves; this file only provides the host-side
// plumbing: allocate exportable buffers, exchange handles (done in Python over the existing process
// group), map peers, and wrap raw pointers as tensors.
#include <cuda_runtime.h>
#include <torch/extension.h>

#include <string>

#include "peer_mem.h"

namespace rfa {
namespace {

void cuda_check(cudaError_t e, const char* what) {
  TORCH_CHECK(e == cudaSuccess, "ring_flash_attn_b200 peer memory: ", what, ": ", cudaGetErrorString(e));
}

// Allocate `bytes` with cudaMalloc (IPC-exportable, unlike the caching allocator's suballocations),
// zero it and return (device pointer, opaque 64-byte IPC handle).
pybind11::tuple peer_alloc(int64_t bytes, int64_t device) {
  int prev = 0;
  cuda_check(cudaGetDevice(&prev), "cudaGetDevice");
  cuda_check(cudaSetDevice(static_cast<int>(device)), "cudaSetDevice");
  void* ptr = nullptr;
  cuda_check(cudaMalloc(&ptr, static_cast<size_t>(bytes)), "cudaMalloc");
  cuda_check(cudaMemset(ptr, 0, static_cast<size_t>(bytes)), "cudaMemset");
  cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  cudaIpcMemHandle_t h;
  cuda_check(cudaIpcGetMemHandle(&h, ptr), "cudaIpcGetMemHandle");
  cuda_check(cudaSetDevice(prev), "cudaSetDevice");
  return pybind11::make_tuple(reinterpret_cast<int64_t>(ptr),
                              pybind11::bytes(reinterpret_cast<const char*>(&h), sizeof(h)));
}

int64_t peer_open(const std::string& handle, int64_t device) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  int prev = 0;
  cuda_check(cudaGetDevice(&prev), "cudaGetDevice");
  cuda_check(cudaSetDevice(static_cast<int>(device)), "cudaSetDevice");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle.data(), sizeof(h));
  void* ptr = nullptr;
  cuda_check(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  cuda_check(cudaSetDevice(prev), "cudaSetDevice");
  return reinterpret_cast<int64_t>(ptr);
}

void peer_close(int64_t ptr) { cuda_check(cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)), "cudaIpcCloseMemHandle"); }
void peer_free(int64_t ptr) { cuda_check(cudaFree(reinterpret_cast<void*>(ptr)), "cudaFree"); }

bool can_access_peer(int64_t device, int64_t peer) {
  int ok = 0;
  cuda_check(cudaDeviceCanAccessPeer(&ok, static_cast<int>(device), static_cast<int>(peer)), "cudaDeviceCanAccessPeer");
  return ok != 0;
}

at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, at::ScalarType dtype, int64_t device) {
  auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, static_cast<c10::DeviceIndex>(device));
  return at::from_blob(reinterpret_cast<void*>(ptr), sizes, opts);
}

}  // namespace

void bind_peer_mem(pybind11::module_& m) {
  m.def("peer_alloc", &peer_alloc, "cudaMalloc + IPC handle: (bytes, device) -> (ptr, handle)");
  m.def("peer_open", &peer_open, "map a peer's IPC handle: (handle, device) -> ptr");
  m.def("peer_close", &peer_close);
  m.def("peer_free", &peer_free);
  m.def("can_access_peer", &can_access_peer);
  m.def("tensor_from_ptr", &tensor_from_ptr, "wrap a raw device pointer as a tensor (no ownership)");
}

}  // namespace rfa
