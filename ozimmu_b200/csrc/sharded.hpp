// sharded.hpp -- one NCCL communicator as the multi-GPU paths use it (sharded.cpp).
//
// The reference has no multi-GPU path.  The partition is the one BASELINE.json names: GPU g owns a row
// block of A and C, B is replicated by a broadcast from its owner over NVLink 5 / NVSwitch, and every GPU runs the
// single-GPU path on its block -- bit-identical to one GPU, because the split scales A per row and B per column and K
// is never partitioned (no reduction between shards).  One process per GPU.
#pragma once
#include <cstddef>
#include <vector>

#include <cuda_runtime.h>

namespace oz {
namespace host {

struct Comm {
  void *nccl = nullptr;          // ncclComm_t
  bool owned = false;            // created by ozimmu_comm_create (destroyed with the Comm) or adopted from the caller
  int rank = 0, size = 1, device = 0;
  cudaStream_t stream = nullptr; // the collectives of a call are enqueued here, ordered against the caller's stream by events
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev_panel;  // "panel p of B has arrived", grown on demand
};

// ncclBroadcast of `count` doubles in place (buf is the send buffer on `root`, the receive buffer elsewhere) on stream
// `s`; throws std::runtime_error on failure
void comm_broadcast_f64(Comm *c, double *buf, std::size_t count, int root, cudaStream_t s);
cudaEvent_t comm_panel_event(Comm *c, std::size_t p);

}  // namespace host
}  // namespace oz
