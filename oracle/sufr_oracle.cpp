// oracle/sufr_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  This is a CPU restatement (C++17, no CUDA) of the reference's
// `create` hot path, used as the parity checker for the CUDA product in `sufr_b200/`.
// Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
// legs may load it.  The product path never calls into this file.
//
// What it restates (reference = TravisWheelerLab/sufr v0.7.12, read-only at /root/reference):
//   libsufr/src/sufr_builder.rs:143-220   SufrBuilder::new        -> Builder::Builder / build()
//   libsufr/src/sufr_builder.rs:241-254   find_n_run              -> Builder::find_n_run
//   libsufr/src/sufr_builder.rs:268-334   find_lcp                -> Builder::find_lcp
//   libsufr/src/sufr_builder.rs:346-382   is_less                 -> Builder::is_less
//   libsufr/src/sufr_builder.rs:391-394   upper_bound             -> Builder::upper_bound
//   libsufr/src/sufr_builder.rs:404-487   partition               -> Builder::partition
//   libsufr/src/sufr_builder.rs:495-598   sort                    -> Builder::sort
//   libsufr/src/sufr_builder.rs:601-767   merge_sort / merge      -> Builder::merge_sort / merge
//   libsufr/src/sufr_builder.rs:771-809   select_pivots           -> Builder::select_pivots
//   libsufr/src/sufr_builder.rs:817-918   write                   -> Builder::serialize
//   libsufr/src/types.rs:80-200           SeedMask                -> struct SeedMask
//   libsufr/src/util.rs:19-37             find_lcp_full_offset    -> full_offset
//   libsufr/src/util.rs:51-89             read_sequence_file      -> read_sequence_file
//   libsufr/src/suffix_array.rs:460-470   u32/u64 dispatch        -> oracle_build(index_bits = 0)
//
// Pinning: tests/test_oracle_golden.py checks this restatement byte-for-byte against the 14
// version-6 golden `.sufr` files of the reference (copied to tests/golden/expected) and the
// in-source known-answer vectors of libsufr/src/{lib,sufr_builder,util,types}.rs.
//
// Parity UNPINNED for one third-party piece: the pivot positions.  The reference draws them with
// `rand` 0.9 `StdRng::seed_from_u64` + `random_range` (sufr_builder.rs:781-788); that crate is not
// vendored under /root/reference and no reference test pins a pivot value.  We draw pivots with
// splitmix64 instead.  In the full-sort and seed-mask modes the output is independent of the pivots
// (SURVEY.md section 8a "Semantics distilled"), which the tests verify by varying seed / partitions.
//
// Differences from the reference that do not change results: partitions live in RAM instead of
// temp files; threads are std::thread over contiguous position ranges (per-partition input order is
// ascending position, i.e. the reference's `-t 1` order).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <string>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>
#include <chrono>

namespace {

// ---------------------------------------------------------------- types.rs:36-200
struct SeedMask {
    std::string mask;
    std::vector<uint8_t> bytes;
    std::vector<size_t> positions;
    std::vector<size_t> differences;
    size_t weight = 0;

    // types.rs:163-166  regex ^1+0[01]*1$
    static bool is_valid(const std::string& m) {
        // equivalent to: only 0/1, first char 1, last char 1, at least one 0
        size_t n = m.size();
        if (n < 3 || m[0] != '1' || m[n - 1] != '1') return false;
        bool zero = false;
        for (char c : m) {
            if (c == '0') zero = true;
            else if (c != '1') return false;
        }
        return zero;
    }
    static bool make(const std::string& m, SeedMask& out) {
        if (!is_valid(m)) return false;
        out.mask = m;
        out.bytes.clear();
        for (char c : m) out.bytes.push_back(c == '1' ? 1 : 0);  // types.rs:170-179
        out.positions.clear();
        for (size_t i = 0; i < out.bytes.size(); i++)
            if (out.bytes[i] == 1) out.positions.push_back(i);  // types.rs:192-199
        out.differences.clear();
        for (size_t i = 0; i < out.positions.size(); i++)
            out.differences.push_back(out.positions[i] - i);  // types.rs:183-189
        out.weight = out.positions.size();
        return true;
    }
};

// ---------------------------------------------------------------- util.rs:19-37
size_t full_offset(size_t lcp, const SeedMask* mask) {
    if (!mask) return lcp;
    if (lcp == 0 || lcp > mask->bytes.size()) return lcp;
    size_t offset = mask->positions[lcp - 1];
    size_t next_offset = lcp < mask->positions.size() ? mask->positions[lcp] : 0;
    if (next_offset > offset && next_offset - offset > 1) return next_offset;
    return offset + 1;
}

uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct PartitionInfo {  // sufr_builder.rs:924-942
    size_t order, len, first_suffix, last_suffix, offset;
};

struct PhaseTimes {
    double transform_s = 0, nscan_s = 0, pivots_s = 0, partition_s = 0, sort_s = 0, stitch_s = 0;
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- sufr_builder.rs:38-89
template <typename T>
struct Builder {
    bool is_dna = false, allow_ambiguity = false, ignore_softmask = false;
    std::vector<uint8_t> text;
    size_t text_len = 0;
    bool has_mask = false;
    SeedMask mask;
    size_t max_query_len = 0;  // SuffixSortType::MaxQueryLen(n); 0 = full
    std::vector<std::pair<size_t, size_t>> n_ranges;  // [start, end)
    std::vector<PartitionInfo> partitions;
    std::vector<T> sa, lcp;  // concatenated per-partition results (seam LCP repaired in stitch())
    size_t num_suffixes = 0;
    int threads = 1;
    PhaseTimes times;
    std::string error;
    mutable bool pivot_error = false;

    const SeedMask* mptr() const { return has_mask ? &mask : nullptr; }

    // sufr_builder.rs:143-195: text transform + N-run scan
    void init_text(const uint8_t* t, size_t n) {
        double t0 = now_s();
        text.resize(n);
        for (size_t i = 0; i < n; i++) {
            uint8_t b = t[i];
            if (b >= 97 && b <= 122) b = ignore_softmask ? (uint8_t)'N' : (uint8_t)(b & 0x5F);
            text[i] = b;
        }
        text_len = n;
        times.transform_s = now_s() - t0;
        t0 = now_s();
        n_ranges.clear();
        if (allow_ambiguity) {
            const size_t min_n = 1000;
            bool in_run = false;
            size_t start = 0;
            for (size_t i = 0; i < n; i++) {
                if (text[i] == 'N') {
                    if (!in_run) { in_run = true; start = i; }
                } else {
                    if (in_run && i - start >= min_n) n_ranges.push_back({start, i});
                    in_run = false;
                }
            }
        }
        times.nscan_s = now_s() - t0;
    }

    // sufr_builder.rs:241-254
    bool find_n_run(size_t suffix, size_t& end) const {
        size_t lo = 0, hi = n_ranges.size();
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            const auto& r = n_ranges[mid];
            if (r.first <= suffix && suffix < r.second) { end = r.second; return true; }
            if (r.first < suffix) lo = mid + 1; else hi = mid;
        }
        return false;
    }

    // sufr_builder.rs:268-334
    size_t find_lcp(size_t start1, size_t start2, size_t len, size_t skip) const {
        if (has_mask) {
            size_t count = 0;
            for (size_t k = skip; k < mask.positions.size(); k++) {
                size_t a = start1 + mask.positions[k], b = start2 + mask.positions[k];
                if (a >= text_len || b >= text_len) break;  // filter(<text_len) then zip
                if (text[a] != text[b]) break;
                count++;
            }
            return skip + count;
        }
        size_t end1, end2;
        if (find_n_run(start1, end1) && find_n_run(start2, end2))
            return std::min(end1 - start1, end2 - start2);
        if (max_query_len > 0) len = max_query_len;
        size_t s1 = start1 + skip, s2 = start2 + skip;
        size_t e1 = std::min(s1 + len, text_len), e2 = std::min(s2 + len, text_len);
        size_t count = 0;
        while (s1 + count < e1 && s2 + count < e2 && text[s1 + count] == text[s2 + count]) count++;
        return skip + count;
    }

    // sufr_builder.rs:346-382
    bool is_less(size_t start1, size_t start2) const {
        if (start1 == start2) return false;
        size_t mq = has_mask ? mask.weight : (max_query_len > 0 ? max_query_len : text_len);
        size_t len_lcp = full_offset(find_lcp(start1, start2, mq, 0), mptr());
        if (len_lcp >= mq) return false;
        bool ha = start1 + len_lcp < text_len, hb = start2 + len_lcp < text_len;
        if (ha && hb) return text[start1 + len_lcp] < text[start2 + len_lcp];
        if (!ha && hb) return true;
        return false;
    }

    // sufr_builder.rs:391-394  (slice::partition_point)
    size_t upper_bound(size_t suffix, const T* pivots, size_t np) const {
        size_t lo = 0, hi = np;
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (is_less((size_t)pivots[mid], suffix)) lo = mid + 1; else hi = mid;
        }
        return lo;
    }

    // sufr_builder.rs:634-767
    void merge(T* suffix_array, size_t n, size_t mid, T* lcp_w, T* target_sa, T* target_lcp) const {
        T* x = suffix_array; T* y = suffix_array + mid;
        T* lcp_x = lcp_w;    T* lcp_y = lcp_w + mid;
        size_t len_x = mid, len_y = n - mid;
        size_t m = 0, idx_x = 0, idx_y = 0, idx_t = 0;
        while (idx_x < len_x && idx_y < len_y) {
            size_t l_x = (size_t)lcp_x[idx_x];
            size_t len_lcp_store = 0;
            if (l_x > m) {
                target_sa[idx_t] = x[idx_x];
                target_lcp[idx_t] = (T)l_x;
            } else if (l_x < m) {
                target_sa[idx_t] = y[idx_y];
                target_lcp[idx_t] = (T)m;
                m = l_x;
            } else {
                size_t xs = (size_t)x[idx_x], ys = (size_t)y[idx_y];
                size_t shorter = std::max(xs, ys);
                size_t max_n = text_len - shorter;
                size_t context;
                if (has_mask) {
                    context = 0;
                    for (size_t p : mask.positions) if (p < max_n) context++;
                } else if (max_query_len > 0) {
                    context = std::min(max_query_len, max_n);
                } else {
                    context = max_n;
                }
                size_t len_lcp, full_len;
                if (m < context) {
                    len_lcp = find_lcp(xs, ys, context - m, m);
                    full_len = full_offset(len_lcp, mptr());
                } else {
                    len_lcp = context; full_len = context;
                }
                if (len_lcp >= context) {
                    target_sa[idx_t] = (T)shorter;
                } else {
                    uint8_t a = text[xs + full_len], b = text[ys + full_len];
                    if (a == b) target_sa[idx_t] = (T)shorter;
                    else if (a < b) target_sa[idx_t] = x[idx_x];
                    else target_sa[idx_t] = y[idx_y];
                }
                if (target_sa[idx_t] == x[idx_x]) target_lcp[idx_t] = (T)l_x;
                else target_lcp[idx_t] = (T)m;
                m = len_lcp;
                (void)len_lcp_store;
            }
            if (target_sa[idx_t] == x[idx_x]) {
                idx_x++;
            } else {
                idx_y++;
                std::swap(x, y);
                std::swap(len_x, len_y);
                std::swap(lcp_x, lcp_y);
                std::swap(idx_x, idx_y);
            }
            idx_t++;
        }
        while (idx_x < len_x) {
            target_sa[idx_t] = x[idx_x];
            target_lcp[idx_t] = lcp_x[idx_x];
            idx_x++; idx_t++;
        }
        if (idx_y < len_y) {
            target_sa[idx_t] = y[idx_y];
            target_lcp[idx_t] = (T)m;
            idx_y++; idx_t++;
            while (idx_y < len_y) {
                target_sa[idx_t] = y[idx_y];
                target_lcp[idx_t] = lcp_y[idx_y];
                idx_y++; idx_t++;
            }
        }
    }

    // sufr_builder.rs:601-631
    void merge_sort(T* x, T* y, size_t n, T* lcp, T* lcp_w) const {
        if (n == 1) {
            lcp[0] = 0;
        } else {
            size_t mid = n / 2;
            merge_sort(y, x, mid, lcp_w, lcp);
            merge_sort(y + mid, x + mid, n - mid, lcp_w + mid, lcp + mid);
            merge(x, n, mid, lcp_w, y, lcp);
        }
    }

    // sufr_builder.rs:771-809 (RNG replaced by splitmix64: pivot values are UNPINNED, see header)
    std::vector<T> select_pivots(size_t num_partitions, uint64_t random_seed) const {
        std::vector<T> pivots;
        if (num_partitions <= 1) return pivots;
        size_t num_pivots = num_partitions - 1;
        if (is_dna) {
            // The reference draws positions until it has num_pivots distinct ACGT$ ones and loops forever when
            // the text has fewer (sufr_builder.rs:787-796).  The oracle reports that case instead of hanging.
            size_t eligible = 0;
            for (size_t i = 0; i < text_len && eligible < num_pivots; i++) {
                uint8_t c = text[i];
                if (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '$') eligible++;
            }
            if (eligible < num_pivots) {
                pivot_error = true;
                return pivots;
            }
        }
        uint64_t state = random_seed > 0 ? random_seed : 0x5DEECE66Dull ^ (uint64_t)text_len;
        std::unordered_set<uint64_t> seen;
        std::vector<uint64_t> order;
        while (true) {
            size_t pos = (size_t)(splitmix64(state) % (uint64_t)text_len);
            uint8_t c = text[pos];
            if (is_dna && !(c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == '$')) continue;
            if (seen.insert(pos).second) order.push_back(pos);
            if (seen.size() == num_pivots) break;
        }
        std::sort(order.begin(), order.end());  // HashSet iteration order is arbitrary in the reference
        pivots.assign(order.begin(), order.end());
        std::vector<T> sa_w(pivots), l(num_pivots, 0), lw(num_pivots, 0);
        merge_sort(sa_w.data(), pivots.data(), num_pivots, l.data(), lw.data());
        return pivots;
    }

    static bool indexed_byte(uint8_t v, bool is_dna, bool allow_ambiguity) {  // sufr_builder.rs:446-449
        return v == '$' || !is_dna || (v == 'A' || v == 'C' || v == 'G' || v == 'T') || allow_ambiguity;
    }

    // sufr_builder.rs:404-487 + 495-598
    bool sort(size_t num_partitions, uint64_t random_seed) {
        // -- partition(): inflate the partition count (sufr_builder.rs:411-422)
        size_t max_partitions = text_len / 4;
        size_t raw_parts;
        if (num_partitions * 10 < max_partitions) raw_parts = num_partitions * 10;
        else if (num_partitions * 5 < max_partitions) raw_parts = num_partitions * 5;
        else if (num_partitions * 2 < max_partitions) raw_parts = num_partitions * 2;
        else if (num_partitions < max_partitions) raw_parts = num_partitions;
        else raw_parts = max_partitions;

        double t0 = now_s();
        std::vector<T> pivots = select_pivots(raw_parts, random_seed);
        times.pivots_s = now_s() - t0;
        if (pivot_error) {
            error = "reference would loop forever: fewer ACGT$ positions than pivots (sufr_builder.rs:787-796)";
            return false;
        }

        // HOT LOOP A (sufr_builder.rs:442-462): every indexed position -> upper_bound -> partition
        t0 = now_s();
        int nt = std::max(1, threads);
        std::vector<std::vector<std::vector<T>>> local(nt, std::vector<std::vector<T>>(std::max<size_t>(raw_parts, 1)));
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; t++) {
                pool.emplace_back([&, t]() {
                    size_t lo = text_len * (size_t)t / nt, hi = text_len * (size_t)(t + 1) / nt;
                    auto& mine = local[t];
                    for (size_t i = lo; i < hi; i++) {
                        if (indexed_byte(text[i], is_dna, allow_ambiguity)) {
                            size_t part = upper_bound(i, pivots.data(), pivots.size());
                            mine[part].push_back((T)i);
                        }
                    }
                });
            }
            for (auto& th : pool) th.join();
        }
        size_t nbuilders = raw_parts;  // `for _ in 0..num_partitions` builders; 0 builders if raw_parts == 0
        std::vector<size_t> total_len(nbuilders, 0);
        size_t total = 0;
        for (size_t p = 0; p < nbuilders; p++) {
            for (int t = 0; t < nt; t++) total_len[p] += local[t][p].size();
            total += total_len[p];
        }
        if (nbuilders == 0) {
            // upper_bound returns 0 with no pivots and builders[0] does not exist: the reference panics.
            for (int t = 0; t < nt; t++) if (!local[t].empty() && !local[t][0].empty()) {
                error = "index out of bounds: no partitions (text too short)";
                return false;
            }
        }
        times.partition_s = now_s() - t0;

        // -- sort(): coalesce raw partitions into <= num_partitions groups (sufr_builder.rs:499-539)
        t0 = now_s();
        if (num_partitions == 0) {
            if (total != 0) { error = "Took 0 but needed to take " + std::to_string(total); return false; }
        }
        size_t num_per_partition = num_partitions ? (size_t)std::ceil((double)total / (double)num_partitions) : 0;
        std::vector<std::vector<size_t>> inputs(num_partitions);
        size_t num_taken = 0, next = 0;
        for (size_t g = 0; g < num_partitions; g++) {
            size_t boundary = num_per_partition * (g + 1);
            while (next < nbuilders) {
                size_t b = next++;
                if (total_len[b] > 0) { inputs[g].push_back(b); num_taken += total_len[b]; }
                if (g + 1 < num_partitions && num_taken > boundary) break;
            }
        }
        if (num_taken != total) {
            error = "Took " + std::to_string(num_taken) + " but needed to take " + std::to_string(total);
            return false;
        }

        // HOT LOOP B (sufr_builder.rs:544-581): one task per coalesced partition
        std::vector<size_t> glen(num_partitions, 0), goff(num_partitions, 0);
        size_t off = 0;
        for (size_t g = 0; g < num_partitions; g++) {
            for (size_t b : inputs[g]) glen[g] += total_len[b];
            goff[g] = off; off += glen[g];
        }
        sa.assign(total, 0);
        lcp.assign(total, 0);
        std::atomic<size_t> next_group{0};
        auto worker = [&]() {
            while (true) {
                size_t g = next_group.fetch_add(1);
                if (g >= num_partitions) break;
                size_t len = glen[g];
                if (len == 0) continue;
                T* part_sa = sa.data() + goff[g];
                size_t w = 0;
                for (size_t b : inputs[g])
                    for (int t = 0; t < nt; t++) {
                        auto& v = local[t][b];
                        std::copy(v.begin(), v.end(), part_sa + w);
                        w += v.size();
                        std::vector<T>().swap(v);
                    }
                std::vector<T> sa_w(part_sa, part_sa + len), lcp_w(len, 0);
                merge_sort(sa_w.data(), part_sa, len, lcp.data() + goff[g], lcp_w.data());
            }
        };
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < nt; t++) pool.emplace_back(worker);
            for (auto& th : pool) th.join();
        }
        partitions.clear();
        for (size_t g = 0; g < num_partitions; g++)
            if (glen[g] > 0)
                partitions.push_back({g, glen[g], (size_t)sa[goff[g]], (size_t)sa[goff[g] + glen[g] - 1], goff[g]});
        num_suffixes = total;
        times.sort_s = now_s() - t0;

        // seam repair, done while stitching in write() (sufr_builder.rs:886-906)
        t0 = now_s();
        for (size_t i = 1; i < partitions.size(); i++)
            lcp[partitions[i].offset] =
                (T)find_lcp(partitions[i - 1].last_suffix, partitions[i].first_suffix, text_len, 0);
        times.stitch_s = now_s() - t0;
        return true;
    }
};

void put_u64(std::vector<uint8_t>& out, uint64_t v) {  // util.rs:138-151 (8 little-endian bytes)
    for (int i = 0; i < 8; i++) out.push_back((uint8_t)(v >> (8 * i)));
}

// sufr_builder.rs:817-918: header + text + SA + LCP + bincode(names)
template <typename T>
std::vector<uint8_t> serialize(const Builder<T>& b, const std::vector<uint64_t>& starts,
                               const std::vector<std::string>& names) {
    std::vector<uint8_t> out;
    out.push_back(6);  // OUTFILE_VERSION types.rs:16
    out.push_back(b.is_dna ? 1 : 0);
    out.push_back(b.allow_ambiguity ? 1 : 0);
    out.push_back(b.ignore_softmask ? 1 : 0);
    put_u64(out, b.text_len);
    size_t locs_pos = out.size();
    put_u64(out, 0); put_u64(out, 0); put_u64(out, 0);
    put_u64(out, b.num_suffixes);
    put_u64(out, b.has_mask ? 0 : b.max_query_len);
    put_u64(out, starts.size());
    for (uint64_t s : starts) {
        T v = (T)s;
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
        out.insert(out.end(), p, p + sizeof(T));
    }
    if (b.has_mask) {
        put_u64(out, b.mask.bytes.size());
        out.insert(out.end(), b.mask.bytes.begin(), b.mask.bytes.end());
    } else {
        put_u64(out, 0);
    }
    uint64_t text_pos = out.size();
    out.insert(out.end(), b.text.begin(), b.text.end());
    uint64_t sa_pos = out.size();
    const uint8_t* sp = reinterpret_cast<const uint8_t*>(b.sa.data());
    out.insert(out.end(), sp, sp + b.sa.size() * sizeof(T));
    uint64_t lcp_pos = out.size();
    const uint8_t* lp = reinterpret_cast<const uint8_t*>(b.lcp.data());
    out.insert(out.end(), lp, lp + b.lcp.size() * sizeof(T));
    put_u64(out, names.size());  // bincode 1.3 Vec<String>: u64 count, then u64 len + bytes
    for (const auto& s : names) {
        put_u64(out, s.size());
        out.insert(out.end(), s.begin(), s.end());
    }
    std::vector<uint8_t> tmp;
    put_u64(tmp, text_pos); put_u64(tmp, sa_pos); put_u64(tmp, lcp_pos);
    std::copy(tmp.begin(), tmp.end(), out.begin() + locs_pos);
    return out;
}

// ---------------------------------------------------------------- util.rs:51-89 (needletail FASTA/FASTQ subset)
struct SeqData {
    std::vector<uint8_t> seq;
    std::vector<uint64_t> starts;
    std::vector<std::string> names;
    std::string error;
};

bool read_sequence_file(const char* path, uint8_t delim, SeqData& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { out.error = std::string(path) + ": cannot open"; return false; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (data.empty()) { out.error = "Failed to read the first two bytes. Is the file empty?"; return false; }
    size_t pos = 0, n = data.size();
    auto next_line = [&](std::string& line) -> bool {
        if (pos >= n) return false;
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) e = n;
        line.assign(data, pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };
    char first = data[0];
    if (first != '>' && first != '@') { out.error = "Bad starting byte; not FASTA/FASTQ"; return false; }
    size_t i = 0;
    std::string line;
    auto push_record = [&](const std::string& header, const std::string& body) {
        if (i > 0) out.seq.push_back(delim);
        out.starts.push_back(out.seq.size());
        out.seq.insert(out.seq.end(), body.begin(), body.end());
        i += 1;
        // id up to first whitespace; fallback (i + 1) with i already incremented (util.rs:73-76)
        size_t s = 0;
        while (s < header.size() && isspace((unsigned char)header[s])) s++;
        size_t e = s;
        while (e < header.size() && !isspace((unsigned char)header[e])) e++;
        out.names.push_back(e > s ? header.substr(s, e - s) : std::to_string(i + 1));
    };
    if (first == '>') {
        std::string header, body;
        bool have = false;
        while (next_line(line)) {
            if (!line.empty() && line[0] == '>') {
                if (have) push_record(header, body);
                header = line.substr(1); body.clear(); have = true;
            } else {
                body += line;
            }
        }
        if (have) push_record(header, body);
    } else {
        std::string header, body, plus, qual;
        while (next_line(header)) {
            if (header.empty()) continue;
            if (!next_line(body) || !next_line(plus) || !next_line(qual)) { out.error = "truncated FASTQ record"; return false; }
            push_record(header.substr(1), body);
        }
    }
    out.seq.push_back('$');  // SENTINEL_CHARACTER types.rs:20
    return true;
}

struct Handle {
    int bits = 32;
    Builder<uint32_t> b32;
    Builder<uint64_t> b64;
    std::vector<uint64_t> starts;
    std::vector<std::string> names;
    std::vector<uint8_t> file_bytes;
    std::vector<uint64_t> n_ranges_flat;
};

}  // namespace

extern "C" struct OracleArgs {
    const uint8_t* text;
    uint64_t text_len;
    int32_t is_dna, allow_ambiguity, ignore_softmask;
    int32_t has_max_query_len;
    uint64_t max_query_len;
    const char* seed_mask;  // NULL = none
    uint64_t num_partitions;
    uint64_t random_seed;
    int32_t threads;
    int32_t index_bits;  // 32, 64, or 0 = reference dispatch (suffix_array.rs:460-470)
    const uint64_t* sequence_starts;
    uint64_t num_sequences;
    const char* const* sequence_names;
};

static void set_err(char* err, size_t errlen, const std::string& msg) {
    if (err && errlen) { snprintf(err, errlen, "%s", msg.c_str()); }
}

template <typename T>
static bool setup(Builder<T>& b, const OracleArgs* a, char* err, size_t errlen) {
    b.is_dna = a->is_dna; b.allow_ambiguity = a->allow_ambiguity; b.ignore_softmask = a->ignore_softmask;
    b.threads = a->threads > 0 ? a->threads : 1;
    b.init_text(a->text, a->text_len);
    if (a->seed_mask && a->has_max_query_len) {  // sufr_builder.rs:163-165
        set_err(err, errlen, "Cannot use max_query_len and seed_mask together");
        return false;
    }
    if (a->seed_mask) {
        if (!SeedMask::make(a->seed_mask, b.mask)) {  // types.rs:81-83
            set_err(err, errlen, std::string("Invalid seed mask '") + a->seed_mask + "'");
            return false;
        }
        b.has_mask = true;
    } else {
        b.max_query_len = a->has_max_query_len ? a->max_query_len : 0;
    }
    return true;
}

extern "C" {

void* oracle_build(const OracleArgs* a, int do_sort, char* err, size_t errlen) {
    Handle* h = new Handle();
    int bits = a->index_bits;
    if (bits == 0) bits = (a->text_len < 0xFFFFFFFFull) ? 32 : 64;  // suffix_array.rs:461
    h->bits = bits;
    for (uint64_t i = 0; i < a->num_sequences; i++) {
        h->starts.push_back(a->sequence_starts[i]);
        h->names.push_back(a->sequence_names ? a->sequence_names[i] : "");
    }
    bool ok;
    if (bits == 32) {
        ok = setup(h->b32, a, err, errlen);
        if (ok && do_sort) { ok = h->b32.sort(a->num_partitions, a->random_seed); if (!ok) set_err(err, errlen, h->b32.error); }
        if (ok && do_sort) h->file_bytes = serialize(h->b32, h->starts, h->names);
        for (auto& r : h->b32.n_ranges) { h->n_ranges_flat.push_back(r.first); h->n_ranges_flat.push_back(r.second); }
    } else {
        ok = setup(h->b64, a, err, errlen);
        if (ok && do_sort) { ok = h->b64.sort(a->num_partitions, a->random_seed); if (!ok) set_err(err, errlen, h->b64.error); }
        if (ok && do_sort) h->file_bytes = serialize(h->b64, h->starts, h->names);
        for (auto& r : h->b64.n_ranges) { h->n_ranges_flat.push_back(r.first); h->n_ranges_flat.push_back(r.second); }
    }
    if (!ok) { delete h; return nullptr; }
    return h;
}

void oracle_free(void* p) { delete static_cast<Handle*>(p); }
int oracle_index_bits(void* p) { return static_cast<Handle*>(p)->bits; }
#define DISPATCH(h, expr32, expr64) (static_cast<Handle*>(h)->bits == 32 ? (expr32) : (expr64))
uint64_t oracle_num_suffixes(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, h->b32.num_suffixes, h->b64.num_suffixes); }
uint64_t oracle_text_len(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, h->b32.text_len, h->b64.text_len); }
const uint8_t* oracle_text(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, h->b32.text.data(), h->b64.text.data()); }
const void* oracle_sa(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, (const void*)h->b32.sa.data(), (const void*)h->b64.sa.data()); }
const void* oracle_lcp(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, (const void*)h->b32.lcp.data(), (const void*)h->b64.lcp.data()); }
uint64_t oracle_num_partitions_built(void* p) { Handle* h = (Handle*)p; return DISPATCH(h, h->b32.partitions.size(), h->b64.partitions.size()); }
uint64_t oracle_num_n_ranges(void* p) { return static_cast<Handle*>(p)->n_ranges_flat.size() / 2; }
const uint64_t* oracle_n_ranges(void* p) { return static_cast<Handle*>(p)->n_ranges_flat.data(); }
const uint8_t* oracle_file_bytes(void* p) { return static_cast<Handle*>(p)->file_bytes.data(); }
uint64_t oracle_file_size(void* p) { return static_cast<Handle*>(p)->file_bytes.size(); }
void oracle_phase_times(void* p, double* out6) {
    Handle* h = (Handle*)p;
    const PhaseTimes& t = DISPATCH(h, h->b32.times, h->b64.times);
    out6[0] = t.transform_s; out6[1] = t.nscan_s; out6[2] = t.pivots_s;
    out6[3] = t.partition_s; out6[4] = t.sort_s; out6[5] = t.stitch_s;
}

// unit-level entry points mirroring the private methods the reference's unit tests call
uint64_t oracle_find_lcp(void* p, uint64_t s1, uint64_t s2, uint64_t len, uint64_t skip) {
    Handle* h = (Handle*)p; return DISPATCH(h, h->b32.find_lcp(s1, s2, len, skip), h->b64.find_lcp(s1, s2, len, skip));
}
int oracle_is_less(void* p, uint64_t s1, uint64_t s2) {
    Handle* h = (Handle*)p; return DISPATCH(h, h->b32.is_less(s1, s2), h->b64.is_less(s1, s2)) ? 1 : 0;
}
uint64_t oracle_upper_bound(void* p, uint64_t suffix, const uint64_t* pivots, uint64_t np) {
    Handle* h = (Handle*)p;
    if (h->bits == 32) {
        std::vector<uint32_t> pv(pivots, pivots + np);
        return h->b32.upper_bound(suffix, pv.data(), np);
    }
    return h->b64.upper_bound(suffix, pivots, np);
}

// SeedMask (types.rs:80-200) and find_lcp_full_offset (util.rs:19-37)
int oracle_seed_mask_valid(const char* m) { return SeedMask::is_valid(m) ? 1 : 0; }
// writes up to cap entries into positions/differences/bytes; returns weight or -1 when invalid
int64_t oracle_seed_mask(const char* m, uint64_t* positions, uint64_t* differences, uint8_t* bytes, uint64_t cap) {
    SeedMask sm;
    if (!SeedMask::make(m, sm)) return -1;
    for (size_t i = 0; i < sm.positions.size() && i < cap; i++) { positions[i] = sm.positions[i]; differences[i] = sm.differences[i]; }
    for (size_t i = 0; i < sm.bytes.size() && i < cap; i++) bytes[i] = sm.bytes[i];
    return (int64_t)sm.weight;
}
uint64_t oracle_find_lcp_full_offset(uint64_t lcp, const char* m) {
    if (!m) return lcp;
    SeedMask sm;
    if (!SeedMask::make(m, sm)) return (uint64_t)-1;
    return full_offset(lcp, &sm);
}

// FASTA/FASTQ ingest (util.rs:51-89)
void* oracle_read_sequence_file(const char* path, uint8_t delim, char* err, size_t errlen) {
    SeqData* d = new SeqData();
    if (!read_sequence_file(path, delim, *d)) { set_err(err, errlen, d->error); delete d; return nullptr; }
    return d;
}
void oracle_seq_free(void* p) { delete static_cast<SeqData*>(p); }
uint64_t oracle_seq_len(void* p) { return static_cast<SeqData*>(p)->seq.size(); }
const uint8_t* oracle_seq_bytes(void* p) { return static_cast<SeqData*>(p)->seq.data(); }
uint64_t oracle_seq_count(void* p) { return static_cast<SeqData*>(p)->starts.size(); }
const uint64_t* oracle_seq_starts(void* p) { return static_cast<SeqData*>(p)->starts.data(); }
const char* oracle_seq_name(void* p, uint64_t i) { return static_cast<SeqData*>(p)->names[i].c_str(); }

}  // extern "C"
