// exchange.cu -- host-side halo plan of a row-partitioned system (no CUDA calls in this file).
//
// What it stands in for: ExchangeCL (parallel/exchange.h:194-243, parallel/exchange.cpp:333-362),
// which DROPS builds from the distributed mesh.  DROPS partitions ELEMENTS (ParMETIS on the
// tetrahedra graph, parallel/partitioner.cpp:251-257) and duplicates the interface DOFs, so its
// vectors come in a "distributed" and an "accumulated" form and every dot product needs an
// accumulation.  Here ROWS are partitioned (owner computes): every global DOF has exactly one
// owner, vectors have one form only, and the only thing a product y = A x needs from the
// neighbours is the set of x entries that the owned rows reference but do not own ("ghosts").
//
// Local numbering of a rank:  [0, n_owned)            owned DOFs, ascending global id
//                             [n_owned, n_owned+n_ghost)  ghosts, grouped by owner rank (ascending),
//                                                         ascending global id inside a group
// so the values received from one neighbour land in one contiguous segment.  The plan is built
// in two steps because the SEND side needs the neighbours' requests:
//   1. exchange_create(owned, ghosts, ghost owners)  -> request(q) = ghost ids owned by q
//   2. the caller transports the requests (torch.distributed in drops_b200/partition.py, any
//      backend) and hands each peer's request to exchange_set_send(q, ids) -> send list
#include "dist.cuh"

#include <algorithm>
#include <numeric>

using namespace db200;

namespace db200 {

int64_t exchange_local_of(const drops_b200_exchange_s* E, int64_t gid)
{
    // owned: binary search over the contiguous ranges
    {
        size_t lo = 0, hi = E->ranges.size();
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (E->ranges[mid].gid0 + E->ranges[mid].len <= gid) lo = mid + 1; else hi = mid;
        }
        if (lo < E->ranges.size() && E->ranges[lo].gid0 <= gid) return E->ranges[lo].local0 + (gid - E->ranges[lo].gid0);
    }
    auto it = std::lower_bound(E->ghost_sorted.begin(), E->ghost_sorted.end(), gid);
    if (it != E->ghost_sorted.end() && *it == gid)
        return (int64_t)E->owned.size() + E->ghost_sorted_slot[(size_t)(it - E->ghost_sorted.begin())];
    return -1;
}

}  // namespace db200

extern "C" {

int drops_b200_exchange_create(int rank, int world, size_t n_owned, const int64_t* owned_gid, size_t n_ghost,
                               const int64_t* ghost_gid, const int32_t* ghost_owner, drops_b200_exchange_t* out)
{
    DB_ARG(out != nullptr, "null out");
    *out = nullptr;
    DB_ARG(world >= 1 && rank >= 0 && rank < world, "rank/world out of range");
    DB_ARG(n_owned == 0 || owned_gid != nullptr, "null owned_gid");
    DB_ARG(n_ghost == 0 || (ghost_gid != nullptr && ghost_owner != nullptr), "null ghost arrays");
    DB_ARG(n_owned + n_ghost < (size_t)2147479552, "owned + ghost DOFs exceed the int32 local index of one partition");
    for (size_t i = 0; i < n_owned; ++i) {
        DB_ARG(owned_gid[i] >= 0, "negative global id");
        DB_ARG(i == 0 || owned_gid[i - 1] < owned_gid[i], "owned_gid must be strictly ascending");
    }
    drops_b200_exchange_s* E = new drops_b200_exchange_s;
    E->rank = rank; E->world = world;
    E->owned.assign(owned_gid, owned_gid + n_owned);
    for (size_t i = 0; i < n_owned;) {             // compress into contiguous ranges
        size_t j = i + 1;
        while (j < n_owned && owned_gid[j] == owned_gid[j - 1] + 1) ++j;
        E->ranges.push_back({owned_gid[i], (int64_t)(j - i), (int64_t)i});
        i = j;
    }
    // ghost slots: (owner, gid) ascending
    std::vector<size_t> order(n_ghost);
    std::iota(order.begin(), order.end(), (size_t)0);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        return ghost_owner[a] != ghost_owner[b] ? ghost_owner[a] < ghost_owner[b] : ghost_gid[a] < ghost_gid[b];
    });
    E->ghost_gid.resize(n_ghost); E->ghost_owner.resize(n_ghost);
    E->recv_off.assign((size_t)world + 1, 0);
    int status = 0;
    for (size_t s = 0; s < n_ghost && status == 0; ++s) {
        const size_t g = order[s];
        const int32_t o = ghost_owner[g];
        if (o < 0 || o >= world || o == rank) { set_error("ghost %lld: owner %d is not another rank", (long long)ghost_gid[g], o); status = DROPS_B200_E_ARG; break; }
        if (s > 0 && ghost_owner[order[s - 1]] == o && ghost_gid[order[s - 1]] == ghost_gid[g]) { set_error("ghost ids must be unique"); status = DROPS_B200_E_ARG; break; }
        E->ghost_gid[s] = ghost_gid[g]; E->ghost_owner[s] = o;
        E->recv_off[(size_t)o + 1] += 1;
    }
    if (status < 0) { delete E; return status; }
    for (int q = 0; q < world; ++q) E->recv_off[(size_t)q + 1] += E->recv_off[(size_t)q];
    // gid-sorted view for lookups
    std::vector<size_t> by_gid(n_ghost);
    std::iota(by_gid.begin(), by_gid.end(), (size_t)0);
    std::sort(by_gid.begin(), by_gid.end(), [&](size_t a, size_t b) { return E->ghost_gid[a] < E->ghost_gid[b]; });
    E->ghost_sorted.resize(n_ghost); E->ghost_sorted_slot.resize(n_ghost);
    for (size_t i = 0; i < n_ghost; ++i) { E->ghost_sorted[i] = E->ghost_gid[by_gid[i]]; E->ghost_sorted_slot[i] = (int32_t)by_gid[i]; }
    for (size_t i = 0; i < n_ghost; ++i) {
        if (i > 0 && E->ghost_sorted[i] == E->ghost_sorted[i - 1]) { set_error("ghost id %lld listed twice", (long long)E->ghost_sorted[i]); delete E; return DROPS_B200_E_ARG; }
        // a ghost must not be owned
        size_t lo = 0, hi = E->ranges.size();
        const int64_t gid = E->ghost_sorted[i];
        while (lo < hi) { const size_t mid = (lo + hi) / 2; if (E->ranges[mid].gid0 + E->ranges[mid].len <= gid) lo = mid + 1; else hi = mid; }
        if (lo < E->ranges.size() && E->ranges[lo].gid0 <= gid) { set_error("ghost id %lld is owned by this rank", (long long)gid); delete E; return DROPS_B200_E_ARG; }
    }
    E->send_idx.assign((size_t)world, std::vector<int32_t>());
    E->send_set.assign((size_t)world, false);
    E->send_set[(size_t)rank] = true;
    *out = E;
    return 0;
}

int drops_b200_exchange_destroy(drops_b200_exchange_t E)
{
    delete E;
    return 0;
}

int drops_b200_exchange_sizes(drops_b200_exchange_t E, size_t* n_owned, size_t* n_ghost, size_t* n_send, int* n_neighbours)
{
    DB_ARG(E != nullptr, "null exchange");
    if (n_owned) *n_owned = E->owned.size();
    if (n_ghost) *n_ghost = E->ghost_gid.size();
    size_t ns = 0; int nn = 0;
    for (int q = 0; q < E->world; ++q) {
        ns += E->send_idx[(size_t)q].size();
        if (!E->send_idx[(size_t)q].empty() || E->recv_off[(size_t)q + 1] > E->recv_off[(size_t)q]) ++nn;
    }
    if (n_send) *n_send = ns;
    if (n_neighbours) *n_neighbours = nn;
    return 0;
}

int drops_b200_exchange_request(drops_b200_exchange_t E, int peer, size_t* count, int64_t* gids)
{
    DB_ARG(E != nullptr && count != nullptr, "null argument");
    DB_ARG(peer >= 0 && peer < E->world, "peer out of range");
    const size_t b = E->recv_off[(size_t)peer], e = E->recv_off[(size_t)peer + 1];
    *count = e - b;
    if (gids) std::copy(E->ghost_gid.begin() + (ptrdiff_t)b, E->ghost_gid.begin() + (ptrdiff_t)e, gids);
    return 0;
}

int drops_b200_exchange_set_send(drops_b200_exchange_t E, int peer, size_t count, const int64_t* gids)
{
    DB_ARG(E != nullptr, "null exchange");
    DB_ARG(peer >= 0 && peer < E->world && peer != E->rank, "peer out of range");
    DB_ARG(count == 0 || gids != nullptr, "null gids");
    std::vector<int32_t> idx(count);
    for (size_t i = 0; i < count; ++i) {
        const int64_t l = exchange_local_of(E, gids[i]);
        if (l < 0 || l >= (int64_t)E->owned.size()) {
            set_error("rank %d asked rank %d for global id %lld, which rank %d does not own", peer, E->rank, (long long)gids[i], E->rank);
            return DROPS_B200_E_ARG;
        }
        idx[i] = (int32_t)l;
    }
    E->send_idx[(size_t)peer].swap(idx);
    E->send_set[(size_t)peer] = true;
    return 0;
}

int drops_b200_exchange_send_list(drops_b200_exchange_t E, int peer, size_t* count, int64_t* local_idx)
{
    DB_ARG(E != nullptr && count != nullptr, "null argument");
    DB_ARG(peer >= 0 && peer < E->world, "peer out of range");
    const std::vector<int32_t>& v = E->send_idx[(size_t)peer];
    *count = v.size();
    if (local_idx) for (size_t i = 0; i < v.size(); ++i) local_idx[i] = v[i];
    return 0;
}

int drops_b200_exchange_local_index(drops_b200_exchange_t E, size_t count, const int64_t* gids, int64_t* local)
{
    DB_ARG(E != nullptr && (count == 0 || (gids != nullptr && local != nullptr)), "null argument");
    for (size_t i = 0; i < count; ++i) local[i] = exchange_local_of(E, gids[i]);
    return 0;
}

int drops_b200_exchange_ready(drops_b200_exchange_t E)
{
    DB_ARG(E != nullptr, "null exchange");
    for (int q = 0; q < E->world; ++q)
        if (!E->send_set[(size_t)q]) return 0;
    return 1;
}

// Ghost discovery on the host: the sorted unique column ids of the local rows that are not owned.
int drops_b200_find_ghosts(size_t n_owned, const int64_t* owned_gid, size_t nnz, const size_t* colind,
                           int64_t* ghost_out, size_t cap, size_t* n_ghost)
{
    DB_ARG(n_ghost != nullptr && (n_owned == 0 || owned_gid != nullptr) && (nnz == 0 || colind != nullptr), "null argument");
    std::vector<int64_t> g;
    for (size_t k = 0; k < nnz; ++k) {
        const int64_t c = (int64_t)colind[k];
        if (!std::binary_search(owned_gid, owned_gid + n_owned, c)) g.push_back(c);
    }
    std::sort(g.begin(), g.end());
    g.erase(std::unique(g.begin(), g.end()), g.end());
    *n_ghost = g.size();
    if (ghost_out) std::copy(g.begin(), g.begin() + (ptrdiff_t)std::min(cap, g.size()), ghost_out);
    return 0;
}

// Row partitioner standing in for ParMETIS (parallel/partitioner.cpp:198-257) on the MATRIX graph:
// breadth-first level structure from a pseudo-peripheral row (as in Cuthill-McKee), rows taken in
// level order and cut into `nparts` consecutive chunks of equal stored-entry count.  Deterministic;
// on mesh graphs the chunks are connected slabs with interfaces of about one level.
int drops_b200_partition_rows(size_t rows, const size_t* rowbeg, const size_t* colind, int nparts, int32_t* part)
{
    DB_ARG(rowbeg != nullptr && part != nullptr && nparts >= 1, "bad argument");
    DB_ARG(rows == 0 || rowbeg[rows] == 0 || colind != nullptr, "null colind");
    std::vector<int64_t> order; order.reserve(rows);
    std::vector<char> seen(rows, 0);
    auto bfs = [&](size_t root, std::vector<int64_t>& out, std::vector<char>& mark) -> size_t {
        const size_t first = out.size();
        out.push_back((int64_t)root); mark[root] = 1;
        size_t last_level_begin = first, level_end = out.size(), head = first;
        while (head < out.size()) {
            if (head == level_end) { last_level_begin = level_end; level_end = out.size(); }
            const size_t r = (size_t)out[head++];
            for (size_t k = rowbeg[r]; k < rowbeg[r + 1]; ++k) {
                const size_t c = colind[k];
                if (c < rows && !mark[c]) { mark[c] = 1; out.push_back((int64_t)c); }
            }
        }
        return (size_t)out[last_level_begin];   // a row of the last level: candidate peripheral row
    };
    for (size_t start = 0; start < rows; ++start) {
        if (seen[start]) continue;
        // two sweeps to find a pseudo-peripheral root of this component
        std::vector<int64_t> tmp; std::vector<char> mark(seen);
        size_t root = bfs(start, tmp, mark);
        tmp.clear(); mark = seen;
        root = bfs(root, tmp, mark);
        bfs(root, order, seen);
    }
    const double total = (double)rowbeg[rows] + (double)rows;   // entries + 1 per row (vector work)
    double acc = 0.0;
    for (size_t i = 0; i < rows; ++i) {
        const size_t r = (size_t)order[i];
        int p = (int)(acc * nparts / (total > 0 ? total : 1.0));
        if (p >= nparts) p = nparts - 1;
        part[r] = p;
        acc += (double)(rowbeg[r + 1] - rowbeg[r]) + 1.0;
    }
    return 0;
}

}  // extern "C"
