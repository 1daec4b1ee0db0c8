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
// Lower bound of the adaptive chunk length: 16 for minibatches of up to 8192 rows (their launches
// run the 8-rows-in-flight layout, for which a 16-edge item is two gather iterations), else 8.
// A function of the batch size only, so every world size / epoch mode cuts hub rows identically.
inline uint32_t default_min_chunk(uint32_t batch) { return batch <= 8192 ? 16u : kMinChunk; }

// Partial sums of a split row are folded in two levels: blocks of kFoldBlock consecutive chunks,
// then the block sums.
constexpr uint32_t kFoldBlock = 32;
// partial-sum / counter slots a split row needs: one per chunk plus one per fold block (counters:
// one per block, then one for the second level -- always fewer than the partial slots)
inline uint32_t hub_slots(uint64_t nchunks) {
    const uint64_t nblk = (nchunks + kFoldBlock - 1) / kFoldBlock;
    return (uint32_t)(nchunks + (nblk > 1 ? nblk : 1));
}

struct HostPlan {
    uint64_t nb = 0;
    std::vector<uint64_t> item_ptr;   // nb+1 offsets into items / hub
    std::vector<uint32_t> n_hub;      // hub-chunk items at the front of each minibatch
    std::vector<Item> items;
    std::vector<HubInfo> hub;
    uint32_t max_slots = 0;
    uint64_t total_slots = 0;         // hub partial-sum slots of all minibatches together
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

// Row ownership on a multi-GPU engine.
//   kAssignSlices   : minibatch [blo, bhi) cut into `world` contiguous slices (what an all-gather
//                     of equal-sized blocks needs);
//   kAssignBalanced : longest-processing-time greedy over the minibatch's rows by cost
//                     deg + kRowCost (rows by descending degree, each to the least-loaded rank;
//                     ties by lower vertex id / lower rank), so every rank gets the same number
//                     of pair updates although R-MAT puts its hubs at the low ids.  Used by the
//                     peer-store exchange, which has no contiguity requirement.
// Both are pure functions of (rowptr, batch, world): every rank computes the same partition.
constexpr int kAssignSlices = 0, kAssignBalanced = 1;
// Flag OR-ed into `assign`: schedule the lightest rows (degree 0..3) right after the hub chunks
// instead of last.  On a multi-GPU engine these rows are almost pure peer-store traffic; issuing
// them early lets the NVLink transfer overlap the heavy rows' compute instead of bursting at the
// end of the launch.
constexpr int kOrderLightFirst = 2;
// Flag OR-ed into `assign`: interleave the light rows (degree 0..3) with the heavier ones, two
// items at a time (the two lane groups of a warp keep items of similar length), in proportion to
// their counts.  Light rows are almost pure output traffic; spreading them over the launch keeps
// the multi-GPU row stores below the NVLink egress rate instead of bursting.
constexpr int kOrderInterleave = 4;
// Flag OR-ed into `assign`: plan for the dataflow epoch kernel (one launch per epoch, minibatches
// overlap and wait only for the rows they read).  Every minibatch's item list is padded to a multiple
// of kFlowPad items with inactive items (v = kNoVertex), so the lane groups of one warp always work on
// the same minibatch, and hub partial-sum slots are numbered over the whole epoch (several
// minibatches are in flight at once): max_slots = slots of the epoch.
constexpr int kPlanFlow = 8;
constexpr uint32_t kFlowPad = 4;
constexpr uint32_t kNoVertex = 0xffffffffu;
constexpr uint64_t kRowCost = 6;      // own row + ~5 negatives per vertex

// Rows of minibatch b owned by `rank`, ascending.  world == 1: the whole minibatch.
inline void owned_rows(const uint64_t* rp, uint64_t first_row, uint64_t nrows, uint32_t batch, int rank, int world,
                       int assign, bool walk, uint64_t b, std::vector<uint32_t>& rows) {
    rows.clear();
    const uint64_t blo = first_row + b * batch, bhi = std::min(blo + (uint64_t)batch, first_row + nrows);
    if (world == 1 || (assign & 1) == kAssignSlices) {
        uint64_t lo, hi;
        slice_range(first_row, nrows, batch, rank, world, b, lo, hi);
        rows.reserve(hi - lo);
        for (uint64_t v = lo; v < hi; v++) rows.push_back((uint32_t)v);
        return;
    }
    const uint64_t m = bhi - blo;
    if (walk) {                                   // every row costs the same: deal round-robin
        for (uint64_t k = (uint64_t)rank; k < m; k += (uint64_t)world) rows.push_back((uint32_t)(blo + k));
        return;
    }
    std::vector<uint32_t> order(m);
    for (uint64_t k = 0; k < m; k++) order[k] = (uint32_t)(blo + k);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) {
        return rp[a + 1] - rp[a] > rp[c + 1] - rp[c];
    });
    std::vector<uint64_t> load((size_t)world, 0);
    for (uint64_t k = 0; k < m; k++) {
        int best = 0;
        for (int r = 1; r < world; r++)
            if (load[r] < load[best]) best = r;
        const uint32_t v = order[k];
        load[best] += rp[v + 1] - rp[v] + kRowCost;
        if (best == rank) rows.push_back(v);
    }
    std::sort(rows.begin(), rows.end());
}

// Work plan for rows [first_row, first_row+nrows) cut into minibatches of `batch` rows
// (minibatch b of the range; for epochs first_row = 0, nrows = n).  Within a minibatch:
// hub chunks first (rows with more than `chunk` edges, cut into equal chunks), then the
// remaining rows by descending degree class, so the heaviest items are scheduled first.
// On a multi-GPU engine only the rows of every minibatch that `rank` owns are planned.
// `par` > 0 makes the chunk length adaptive per minibatch: chunk_b = clamp(edges_b / par,
// min_chunk, chunk) with edges_b the edges of the whole minibatch (all ranks), so that a
// minibatch with little work is still cut into enough items to occupy `par` lane groups in one
// wave and its critical path (the longest item) stays short.
inline void build_host_plan(const uint64_t* rp, uint64_t first_row, uint64_t nrows, uint32_t batch,
                            uint32_t chunk, uint32_t par, bool walk, int rank, int world, int assign, HostPlan& out,
                            uint32_t min_chunk = 0) {          // 0 = default_min_chunk(batch); per-engine tuning option
    const uint64_t nb = (nrows + batch - 1) / batch;
    std::vector<uint64_t> item_ptr(nb + 1, 0);
    std::vector<uint32_t> n_hub(nb, 0), n_slots(nb, 0);
    std::vector<uint64_t> chunk_len(nb, chunk);
    std::vector<std::vector<uint32_t>> mine(nb);
    auto nchunks_of = [&](uint64_t deg, uint64_t ch) -> uint64_t {
        return (!walk && deg > ch) ? (deg + ch - 1) / ch : 1;
    };
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t b = 0; b < (int64_t)nb; b++) {
        std::vector<uint32_t>& rows = mine[b];
        owned_rows(rp, first_row, nrows, batch, rank, world, assign, walk, (uint64_t)b, rows);
        // the chunk length is a function of the WHOLE minibatch, not of this rank's share: a hub row
        // is then cut (and its partial sums folded) identically for every world size, which keeps
        // multi-GPU results bit-identical to the single-GPU run
        const uint64_t blo = first_row + (uint64_t)b * batch, bhi = std::min(blo + (uint64_t)batch, first_row + nrows);
        const uint64_t edges = rp[bhi] - rp[blo];
        uint64_t ch = chunk;
        if (par != 0 && bhi > blo) {
            const uint64_t c = (edges + par - 1) / par;
            ch = std::min<uint64_t>(chunk, std::max<uint64_t>(std::min<uint64_t>(min_chunk ? min_chunk : default_min_chunk(batch), chunk), c));
        }
        chunk_len[b] = ch;
        uint64_t cnt = 0, hubs = 0, slots = 0;
        for (uint32_t v : rows) {
            uint64_t c = nchunks_of(rp[v + 1] - rp[v], ch);
            cnt += c;
            if (c > 1) { hubs += c; slots += hub_slots(c); }
        }
        if (assign & kPlanFlow) cnt = (cnt + kFlowPad - 1) / kFlowPad * kFlowPad;
        item_ptr[b + 1] = cnt;
        n_hub[b] = (uint32_t)hubs;
        n_slots[b] = (uint32_t)slots;
    }
    uint32_t max_slots = 0;
    std::vector<uint64_t> slot_base(nb + 1, 0);
    for (uint64_t b = 0; b < nb; b++) {
        item_ptr[b + 1] += item_ptr[b];
        max_slots = std::max(max_slots, n_slots[b]);
        slot_base[b + 1] = slot_base[b] + n_slots[b];
    }
    if (assign & kPlanFlow) max_slots = (uint32_t)std::min<uint64_t>(slot_base[nb], 0xffffffffull);
    out.total_slots = slot_base[nb];
    const uint64_t total = item_ptr[nb];
    std::vector<Item> items(total ? total : 1);
    std::vector<HubInfo> hub(total ? total : 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t b = 0; b < (int64_t)nb; b++) {
        const std::vector<uint32_t>& rows = mine[b];
        Item* it = items.data() + item_ptr[b];
        HubInfo* hb = hub.data() + item_ptr[b];
        const uint64_t ch = chunk_len[b];
        uint64_t k = 0;
        uint32_t slot = (assign & kPlanFlow) ? (uint32_t)slot_base[b] : 0u;
        for (uint32_t v : rows) {                      // hub chunks
            uint64_t deg = rp[v + 1] - rp[v], nc = nchunks_of(deg, ch);
            if (nc <= 1) continue;
            uint64_t base = deg / nc, extra = deg % nc, e0 = rp[v];
            for (uint64_t c = 0; c < nc; c++) {
                uint32_t len = (uint32_t)(base + (c < extra ? 1 : 0));
                it[k] = Item{v, len | kChunkFlag, e0};
                hb[k] = HubInfo{(uint32_t)c, (uint32_t)nc, slot + (uint32_t)c, (uint32_t)deg};
                e0 += len;
                k++;
            }
            slot += hub_slots(nc);
        }
        // remaining rows: counting sort by degree class (0, 1, 2-3, 4-7, ...), descending
        uint64_t cls_cnt[34] = {0};
        auto cls_of = [](uint64_t deg) -> int { return deg == 0 ? 0 : 64 - __builtin_clzll(deg); };
        for (uint32_t v : rows) {
            uint64_t deg = rp[v + 1] - rp[v];
            if (nchunks_of(deg, ch) > 1) continue;
            cls_cnt[std::min(cls_of(deg), 33)]++;
        }
        uint64_t cls_off[34];
        uint64_t off = k;
        if (assign & kOrderLightFirst)
            for (int c = 0; c <= 2; c++) { cls_off[c] = off; off += cls_cnt[c]; }
        for (int c = 33; c >= ((assign & kOrderLightFirst) ? 3 : 0); c--) { cls_off[c] = off; off += cls_cnt[c]; }
        for (uint32_t v : rows) {
            uint64_t deg = rp[v + 1] - rp[v];
            if (nchunks_of(deg, ch) > 1) continue;
            uint64_t pos = cls_off[std::min(cls_of(deg), 33)]++;
            it[pos] = Item{v, (uint32_t)deg, rp[v]};
            hb[pos] = HubInfo{0, 1, 0, (uint32_t)deg};
        }
        if ((assign & kOrderInterleave) && !(assign & kOrderLightFirst)) {
            // it[k .. end) is sorted by descending degree class: heavy part then light part (classes 2,1,0)
            const uint64_t end = item_ptr[b + 1] - item_ptr[b];
            const uint64_t nlight = cls_cnt[0] + cls_cnt[1] + cls_cnt[2];
            const uint64_t nheavy = end - k - nlight;
            if (nlight && nheavy) {
                std::vector<Item> tmp(it + k, it + end);
                uint64_t o = k, hi_i = 0, li = 0;
                while (hi_i < nheavy || li < nlight) {
                    for (int q = 0; q < 2 && hi_i < nheavy; q++) it[o++] = tmp[hi_i++];
                    const uint64_t want = hi_i >= nheavy ? nlight : (hi_i * nlight / nheavy) & ~1ull;
                    while (li < want) it[o++] = tmp[nheavy + li++];
                }
                for (uint64_t q = k; q < end; q++) hb[q] = HubInfo{0, 1, 0, it[q].len};
            }
        }
        if (assign & kPlanFlow) {                      // inactive padding items at the end of the minibatch
            const uint64_t end = item_ptr[b + 1] - item_ptr[b];
            uint64_t used = 0;
            for (uint32_t v : rows) used += nchunks_of(rp[v + 1] - rp[v], ch);
            for (uint64_t q = used; q < end; q++) { it[q] = Item{kNoVertex, 0, 0}; hb[q] = HubInfo{0, 1, 0, 0}; }
        }
        std::vector<uint32_t>().swap(mine[b]);
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
