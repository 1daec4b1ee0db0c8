// force2vec_b200/csrc/f2v_plan.hpp -- host-side work plan of the force step (no CUDA).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace f2v {

constexpr uint32_t kChunkFlag = 0x80000000u;

// 16-byte work item.  len bit 31 set => chunk of a split (hub) row, described by HubInfo.
struct alignas(16) Item {
    uint32_t v;       // vertex id
    uint32_t len;     // edges in this item (| kChunkFlag)
    uint64_t e0;      // first edge (index into colids)
};
struct alignas(16) HubInfo {
    uint32_t chunk;   // index of this chunk within its row
    uint32_t nchunks; // chunks of the row
    uint32_t slot;    // partial-sum row of this chunk (slot - chunk = row's first slot = counter index)
    uint32_t deg;     // full degree of the row
};

constexpr uint32_t kMinChunk = 8;

// partial-sum / counter slots a split row needs: one per chunk (the counter is the first one)
inline uint32_t hub_slots(uint64_t nchunks) { return (uint32_t)nchunks; }

struct HostPlan {
    uint64_t nb = 0;
    std::vector<uint64_t> item_ptr;   // nb+1 offsets into items / hub
    std::vector<uint32_t> n_hub;      // hub-chunk items at the front of each minibatch
    std::vector<Item> items;
    std::vector<HubInfo> hub;
    uint32_t max_slots = 0;
};

// Row slice of minibatch b owned by `rank`: the minibatch [blo, bhi) is cut into `world`
// contiguous slices of batch/world rows (the all-gather exchanges equal-sized slices).
inline void slice_range(uint64_t first_row, uint64_t nrows, uint32_t batch, int rank, int world, uint64_t b,
                        uint64_t& lo, uint64_t& hi) {
    const uint64_t blo = first_row + b * batch, bhi = std::min(blo + (uint64_t)batch, first_row + nrows);
    if (world == 1) { lo = blo; hi = bhi; return; }
    const uint64_t slice = batch / (uint32_t)world;
    lo = std::min(blo + (uint64_t)rank * slice, bhi);
    hi = std::min(lo + slice, bhi);
}

// Work plan for rows [first_row, first_row+nrows) cut into minibatches of `batch` rows
// (minibatch b of the range; for epochs first_row = 0, nrows = n).  Within a minibatch:
// hub chunks first (rows with more than `chunk` edges, cut into equal chunks), then the
// remaining rows by descending degree class, so the heaviest items are scheduled first.
// On a multi-GPU engine only the rank's slice of every minibatch is planned.
// `par` > 0 makes the chunk length adaptive per minibatch: chunk_b = clamp(edges_b / par,
// kMinChunk, chunk), so that a minibatch with little work is still cut into enough items to
// occupy `par` lane groups in one wave and its critical path (the longest item) stays short.
inline void build_host_plan(const uint64_t* rp, uint64_t first_row, uint64_t nrows, uint32_t batch,
                            uint32_t chunk, uint32_t par, bool walk, int rank, int world, HostPlan& out) {
    const uint64_t nb = (nrows + batch - 1) / batch;
    std::vector<uint64_t> item_ptr(nb + 1, 0);
    std::vector<uint32_t> n_hub(nb, 0), n_slots(nb, 0);
    auto my_range = [&](uint64_t b, uint64_t& lo, uint64_t& hi) {
        slice_range(first_row, nrows, batch, rank, world, b, lo, hi);
    };
    auto chunk_of_batch = [&](uint64_t lo, uint64_t hi) -> uint64_t {
        if (par == 0 || hi <= lo) return chunk;
        const uint64_t edges = rp[hi] - rp[lo];
        const uint64_t c = (edges + par - 1) / par;
        return std::min<uint64_t>(chunk, std::max<uint64_t>(std::min<uint64_t>(kMinChunk, chunk), c));
    };
    auto nchunks_of = [&](uint64_t deg, uint64_t ch) -> uint64_t {
        return (!walk && deg > ch) ? (deg + ch - 1) / ch : 1;
    };
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)nb; b++) {
        uint64_t lo, hi, cnt = 0, hubs = 0;
        my_range((uint64_t)b, lo, hi);
        const uint64_t ch = chunk_of_batch(lo, hi);
        uint64_t slots = 0;
        for (uint64_t v = lo; v < hi; v++) {
            uint64_t c = nchunks_of(rp[v + 1] - rp[v], ch);
            cnt += c;
            if (c > 1) { hubs += c; slots += hub_slots(c); }
        }
        item_ptr[b + 1] = cnt;
        n_hub[b] = (uint32_t)hubs;
        n_slots[b] = (uint32_t)slots;
    }
    uint32_t max_slots = 0;
    for (uint64_t b = 0; b < nb; b++) {
        item_ptr[b + 1] += item_ptr[b];
        max_slots = std::max(max_slots, n_slots[b]);
    }
    const uint64_t total = item_ptr[nb];
    std::vector<Item> items(total ? total : 1);
    std::vector<HubInfo> hub(total ? total : 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t b = 0; b < (int64_t)nb; b++) {
        uint64_t lo, hi;
        my_range((uint64_t)b, lo, hi);
        Item* it = items.data() + item_ptr[b];
        HubInfo* hb = hub.data() + item_ptr[b];
        const uint64_t ch = chunk_of_batch(lo, hi);
        uint64_t k = 0;
        uint32_t slot = 0;
        for (uint64_t v = lo; v < hi; v++) {           // hub chunks
            uint64_t deg = rp[v + 1] - rp[v], nc = nchunks_of(deg, ch);
            if (nc <= 1) continue;
            uint64_t base = deg / nc, extra = deg % nc, e0 = rp[v];
            for (uint64_t c = 0; c < nc; c++) {
                uint32_t len = (uint32_t)(base + (c < extra ? 1 : 0));
                it[k] = Item{(uint32_t)v, len | kChunkFlag, e0};
                hb[k] = HubInfo{(uint32_t)c, (uint32_t)nc, slot + (uint32_t)c, (uint32_t)deg};
                e0 += len;
                k++;
            }
            slot += hub_slots(nc);
        }
        // remaining rows: counting sort by degree class (0, 1, 2-3, 4-7, ...), descending
        uint64_t cls_cnt[34] = {0};
        auto cls_of = [](uint64_t deg) -> int { return deg == 0 ? 0 : 64 - __builtin_clzll(deg); };
        for (uint64_t v = lo; v < hi; v++) {
            uint64_t deg = rp[v + 1] - rp[v];
            if (nchunks_of(deg, ch) > 1) continue;
            cls_cnt[std::min(cls_of(deg), 33)]++;
        }
        uint64_t cls_off[34];
        uint64_t off = k;
        for (int c = 33; c >= 0; c--) { cls_off[c] = off; off += cls_cnt[c]; }
        for (uint64_t v = lo; v < hi; v++) {
            uint64_t deg = rp[v + 1] - rp[v];
            if (nchunks_of(deg, ch) > 1) continue;
            uint64_t pos = cls_off[std::min(cls_of(deg), 33)]++;
            it[pos] = Item{(uint32_t)v, (uint32_t)deg, rp[v]};
            hb[pos] = HubInfo{0, 1, 0, (uint32_t)deg};
        }
    }
    out.nb = nb;
    out.item_ptr.swap(item_ptr);
    out.n_hub.swap(n_hub);
    out.items.swap(items);
    out.hub.swap(hub);
    if (total == 0) { out.items.clear(); out.hub.clear(); }
    out.max_slots = max_slots;
}

}  // namespace f2v
