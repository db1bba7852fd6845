// dist.cuh -- row-partitioned systems across GPUs: halo plan, communicator, partition matrix.
#pragma once
#include "common.cuh"

#include <vector>

// Host-side halo plan (exchange.cu).
struct drops_b200_exchange_s {
    int rank = 0, world = 1;
    std::vector<int64_t> owned;              // owned global ids, ascending; local index = position
    struct Range { int64_t gid0, len, local0; };
    std::vector<Range> ranges;               // the owned ids as contiguous runs
    std::vector<int64_t> ghost_gid;          // by ghost slot: (owner, gid) ascending
    std::vector<int32_t> ghost_owner;
    std::vector<size_t> recv_off;            // world+1: slots received from rank q = [recv_off[q], recv_off[q+1])
    std::vector<int64_t> ghost_sorted;       // ghost ids ascending + their slots (lookup)
    std::vector<int32_t> ghost_sorted_slot;
    std::vector<std::vector<int32_t>> send_idx;   // per peer: owned local indices in the peer's request order
    std::vector<bool> send_set;
};

namespace db200 {

int64_t exchange_local_of(const drops_b200_exchange_s* E, int64_t gid);

enum CommKind { COMM_NCCL = 0, COMM_LOCAL = 1 };

// Sorted unique non-owned column ids of device-resident rows (dist.cu).
int find_ghosts_device(size_t n_owned, const int64_t* owned_host_sorted, size_t nnz, const int64_t* d_colgid,
                       std::vector<int64_t>* out);

}  // namespace db200

struct drops_b200_dist_matrix_s;

// One communicator per process.  COMM_NCCL: this process is rank `rank` of `world`, one GPU each,
// ncclSend/ncclRecv for the halo and ncclAllGather for the reduction partials.  COMM_LOCAL: all
// `world` partitions live in this process on this GPU (device-to-device copies); it exists so that
// the whole partitioned path -- plan, remap, interior/boundary split, rank-ordered reductions --
// is exercised by the single-GPU test suite.
struct drops_b200_comm_s {
    int kind = db200::COMM_NCCL;
    int rank = 0, world = 1;
    void* nccl = nullptr;                    // ncclComm_t
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_halo = nullptr, ev_bnd = nullptr;
    db200::dd* gathered = nullptr;           // world * kMaxQ reduction partials, rank-major
    db200::dd* wide = nullptr;               // world * kWideQ partials: the i+1 products of a classical Gram-Schmidt column
    std::vector<drops_b200_dist_matrix_s*> parts;   // COMM_LOCAL: registered partitions by rank
    // Peer-to-peer reductions ("dist.p2p_reduce"): every rank stores its partials straight into all ranks'
    // buffers (CUDA IPC mappings over NVLink) and raises its sequence number there; k_dist_scalars waits for
    // all sequence numbers.  Two parity halves: a rank can be at most one reduction ahead of another.
    int p2p = 0;                             // 1 ready, -1 unavailable (ncclAllGather is used)
    db200::dd* red_buf = nullptr;            // 2 x world x kMaxQ partials, then world flags (one >= 2 MB allocation)
    unsigned* red_flags = nullptr;
    db200::dd** d_peer_red = nullptr;        // world entries: base of every rank's red_buf (mapped)
    unsigned** d_peer_redflag = nullptr;     // world entries: base of every rank's flag array
    unsigned red_seq = 0;
    bool red_last_p2p = false;               // the reduction just gathered went through the peer path
    int* red_fault = nullptr;
    std::vector<void*> ipc_mapped;
};

struct drops_b200_dist_matrix_s {
    drops_b200_comm_s* comm = nullptr;
    int rank = 0, world = 1;
    size_t n_owned = 0, n_ghost = 0, nnz = 0, n_int = 0, n_bnd = 0, n_send = 0;
    db200::SellOp sell_int, sell_bnd;        // rows without / with ghost columns
    double* diag = nullptr;
    bool diag_ok = false;
    double* ghost = nullptr;                 // n_ghost received x entries
    double* sendbuf = nullptr;               // n_send packed x entries
    int32_t* d_send_idx = nullptr;
    std::vector<size_t> send_off, recv_off;  // world+1 each
    uint64_t version = 0;
    size_t device_bytes = 0;
    std::vector<double*> work;               // solver vectors (n_owned)
    void* d_scalars = nullptr;               // db200::SolverScalars
    double* d_hist = nullptr;
    int hist_alloc = 0;
    double* d_small = nullptr;               // GMRES: Hessenberg column, y, basis pointers, scalars
    size_t small_alloc = 0;
    // Row-ordered numbering of the partition ("dist.reorder"): positions of the interior operator, then of
    // the boundary operator; PCG iterates on vectors of npos entries (padding positions hold 0).
    size_t npos = 0, npos_int = 0;           // npos == 0: no row-ordered twins
    int32_t* ro_perm = nullptr;              // position -> owned row (-1 = padding)
    db200::SellOp sell_bnd_ro;               // boundary rows with columns = positions / npos + ghost slot
    double* diag_ro = nullptr;
    uint8_t* diag_ro_idx = nullptr;      // diag_ro as one byte per position + value list ("pcg.diag_index", krylov.cu)
    int diag_ro_count = 0;
    double diag_ro_values[32] = {};
    int32_t* d_send_idx_ro = nullptr;
    bool block_ordered = false;
    // Peer-to-peer halo ("dist.p2p"): the pack kernel stores the boundary entries of x straight into the
    // neighbours' ghost buffers (NVLink peer stores through CUDA IPC mappings; plain pointers for the
    // in-process transport) and raises a sequence flag there; no send buffer, no NCCL call per exchange.
    int p2p = 0;                             // 0 not set up, 1 ready, -1 unavailable (NCCL send/recv is used)
    unsigned p2p_seq = 0;                    // exchanges done through the peer path (same on every rank)
    unsigned* halo_flags = nullptr;          // world entries in MY memory (behind the ghost entries): flags[q] = last sequence pushed by q
    size_t halo_flag_offset = 0;             // byte offset of halo_flags inside the ghost allocation
    unsigned* p2p_ticket = nullptr;
    int* p2p_fault = nullptr;                // a wait timed out
    unsigned long long* wait_stats = nullptr;   // 4 x {ns, count}: halo wait, p.Ap reduction, (rho, residual) reduction, other reductions
    double** d_peer_dst = nullptr;           // world entries: where my entries for q land (q's ghost + offset)
    unsigned** d_peer_flag = nullptr;        // world entries: my slot in q's flag array
    size_t* d_send_off = nullptr;            // world + 1
    unsigned char* d_expect = nullptr;       // world entries: 1 = q sends to me
    std::vector<void*> ipc_mapped;           // pointers opened with cudaIpcOpenMemHandle
};

namespace db200 {
// Builds the partition operator from device-resident rows with global column ids (dist.cu).
int dist_matrix_build(drops_b200_comm_s* comm, drops_b200_exchange_s* E, size_t nnz, const int64_t* d_rowptr,
                      const int64_t* d_colgid, const double* d_val, uint64_t version, drops_b200_dist_matrix_s** out);
}  // namespace db200
