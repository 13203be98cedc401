// Peer-memory runtime: CUDA-IPC backed buffers that every rank of a node maps into its address space, so
// kernels can store straight into (or load from) another GPU's HBM over NVLink.
#pragma once
#include <pybind11/pybind11.h>

namespace rfa {
void bind_peer_mem(pybind11::module_& m);
}
