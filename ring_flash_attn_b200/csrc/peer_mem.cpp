// CUDA-IPC peer buffers (one process per GPU, single node, NVLink/NVSwitch).
//
// The reference moves K/V and dK/dV with NCCL isend/irecv, all_gather and reduce_scatter
// (/root/reference/ring_flash_attn/utils.py:98-168).  Here communication is plain stores/loads on
// peer-mapped memory issued by the attention kernels themselves; this file only provides the host-side
// plumbing: allocate exportable buffers, exchange handles (done in Python over the existing process
// group), map peers, and wrap raw pointers as tensors.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <torch/extension.h>

#include <mutex>
#include <string>
#include <vector>

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

// ---- copy-engine K/V transport --------------------------------------------------------------------------
// The K/V rows a peer needs can also travel by DMA: cudaMemcpy2DAsync on a side stream straight into the peer's
// staging slot (CUDA-IPC mapping), followed by a 4-byte copy that raises the peer's "rows have landed" epoch.  The
// attention kernel is launched WITHOUT push CTAs and waits on exactly the same flags as before.  Copy engines
// reach the NVLink limit without occupying SMs (the in-kernel TMA push needs 24+ SMs and measured 290-480 GB/s);
// stream order guarantees that the flag copy starts only after the row copies have completed.  Staging reuse
// (the peer must have consumed what was pushed two calls ago) is a stream-ordered wait on this rank's own signal
// pad: cuStreamWaitValue32, resolved at run time so that libcuda is not a link dependency.
PFN_cuStreamWaitValue32_v11070 resolve_wait_value() {
  static PFN_cuStreamWaitValue32_v11070 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuStreamWaitValue32_v11070>(p);
  });
  return fn;
}

bool dma_transport_available() { return resolve_wait_value() != nullptr; }

// tasks: CPU int64 (n, 4) rows (src_row, dst_off, rows | dst << 32, which) sorted by destination (the push table of
// parallel/symm.py).  flag_host: pinned int32 ring, flag_dev: device int32 ring (same length), slot = epoch % len.
void kv_push_dma(const at::Tensor& k, const at::Tensor& v, const at::Tensor& tasks, std::vector<int64_t> stage_ptrs,
                 std::vector<int64_t> pad_ptrs, int64_t my_pad_ptr, int64_t parity_off, int64_t row_bytes,
                 int64_t my_rank, int64_t epoch, at::Tensor& flag_host, at::Tensor& flag_dev, int64_t stream_ptr) {
  TORCH_CHECK(tasks.device().is_cpu() && tasks.scalar_type() == at::kLong && tasks.dim() == 2 && tasks.size(1) == 4 &&
              tasks.is_contiguous());
  TORCH_CHECK(flag_host.is_pinned() && flag_host.scalar_type() == at::kInt && flag_dev.scalar_type() == at::kInt &&
              flag_dev.is_cuda() && flag_host.numel() == flag_dev.numel());
  auto wait_value = resolve_wait_value();
  TORCH_CHECK(wait_value != nullptr, "cuStreamWaitValue32 is not available from this driver");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_ptr);
  const int64_t slot = epoch % flag_host.numel();
  flag_host.data_ptr<int>()[slot] = static_cast<int>(epoch);
  int* flag_src = flag_dev.data_ptr<int>() + slot;
  cuda_check(cudaMemcpyAsync(flag_src, flag_host.data_ptr<int>() + slot, 4, cudaMemcpyHostToDevice, stream),
             "flag staging copy");
  constexpr int kPadKvReady = 0, kPadConsumed = 64;  // csrc/attn_common.h
  const char* src_base[2] = {static_cast<const char*>(k.data_ptr()), static_cast<const char*>(v.data_ptr())};
  const int64_t pitch[2] = {k.stride(0) * k.element_size(), v.stride(0) * v.element_size()};
  const int64_t* t = tasks.data_ptr<int64_t>();
  const int64_t n = tasks.size(0);
  int cur = -1;
  auto raise = [&](int dst) {
    int* peer_flag = reinterpret_cast<int*>(pad_ptrs[dst]) + kPadKvReady + my_rank;
    cuda_check(cudaMemcpyAsync(peer_flag, flag_src, 4, cudaMemcpyDeviceToDevice, stream), "flag copy to the peer");
  };
  for (int64_t i = 0; i < n; ++i) {
    const int64_t src_row = t[4 * i], dst_off = t[4 * i + 1], packed = t[4 * i + 2], which = t[4 * i + 3];
    const int rows = static_cast<int>(packed & 0xffffffff), dst = static_cast<int>(packed >> 32);
    if (dst != cur) {
      if (cur >= 0) raise(cur);
      cur = dst;
      if (epoch > 2) {  // the destination must have finished reading what we pushed into this parity two calls ago
        CUresult r = wait_value(reinterpret_cast<CUstream>(stream),
                                static_cast<CUdeviceptr>(my_pad_ptr + 4 * (kPadConsumed + dst)),
                                static_cast<cuuint32_t>(epoch - 2), CU_STREAM_WAIT_VALUE_GEQ);
        TORCH_CHECK(r == CUDA_SUCCESS, "cuStreamWaitValue32 failed: ", static_cast<int>(r));
      }
    }
    char* d = reinterpret_cast<char*>(stage_ptrs[dst]) + parity_off + dst_off;
    const char* sp = src_base[which] + src_row * pitch[which];
    cuda_check(cudaMemcpy2DAsync(d, static_cast<size_t>(row_bytes), sp, static_cast<size_t>(pitch[which]),
                                 static_cast<size_t>(row_bytes), static_cast<size_t>(rows), cudaMemcpyDeviceToDevice,
                                 stream),
               "cudaMemcpy2DAsync to the peer's staging buffer");
  }
  if (cur >= 0) raise(cur);
}

}  // namespace

void bind_peer_mem(pybind11::module_& m) {
  m.def("dma_transport_available", &dma_transport_available);
  m.def("kv_push_dma", &kv_push_dma, "K/V rows to the peers' staging buffers with the copy engines");
  m.def("peer_alloc", &peer_alloc, "cudaMalloc + IPC handle: (bytes, device) -> (ptr, handle)");
  m.def("peer_open", &peer_open, "map a peer's IPC handle: (handle, device) -> ptr");
  m.def("peer_close", &peer_close);
  m.def("peer_free", &peer_free);
  m.def("can_access_peer", &can_access_peer);
  m.def("tensor_from_ptr", &tensor_from_ptr, "wrap a raw device pointer as a tensor (no ownership)");
}

}  // namespace rfa
