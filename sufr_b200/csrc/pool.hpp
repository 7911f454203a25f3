// Device memory pool: a few large cudaMalloc'd chunks carved by a first-fit free list, so that a build
// (and repeated builds on the same context) never calls cudaMalloc/cudaFree in its steady state.
// Sized for B200's 180 GB: one chunk normally holds the whole working set of a build.
#pragma once
#include <algorithm>
#include <map>
#include <vector>

#include "common.cuh"

namespace sufr {

class DevicePool {
   public:
    static constexpr size_t kAlign = 512;

    ~DevicePool() { release_all(); }

    // Make sure one chunk has at least `bytes` of contiguous free space (allocating a new chunk if needed).
    void reserve(size_t bytes) {
        bytes = round_up(bytes);
        for (auto& c : chunks_)
            for (auto& f : c.free_list)
                if (f.second >= bytes) return;
        trim();  // nothing fits: give unused chunks back before asking for a bigger one
        add_chunk(bytes);
    }

    // How many allocations of `bytes` the free lists can serve without growing the pool.
    size_t count_fits(size_t bytes) const {
        bytes = round_up(bytes ? bytes : 1);
        size_t k = 0;
        for (auto& c : chunks_)
            for (auto& f : c.free_list) k += f.second / bytes;
        return k;
    }

    void* alloc(size_t bytes) {
        bytes = round_up(bytes ? bytes : 1);
        for (int attempt = 0; attempt < 2; attempt++) {
            for (size_t ci = 0; ci < chunks_.size(); ci++) {
                auto& c = chunks_[ci];
                for (auto it = c.free_list.begin(); it != c.free_list.end(); ++it) {
                    if (it->second >= bytes) {
                        size_t off = it->first, len = it->second;
                        c.free_list.erase(it);
                        if (len > bytes) c.free_list[off + bytes] = len - bytes;
                        char* p = c.base + off;
                        live_[p] = {ci, bytes};
                        in_use_ += bytes;
                        peak_ = std::max(peak_, in_use_);
                        return p;
                    }
                }
            }
            add_chunk(std::max(bytes, (size_t)2 << 30));  // grow in >= 2 GB steps
        }
        throw Error(3, "device pool: allocation of " + std::to_string(bytes) + " bytes failed");
    }

    void free(void* ptr) {
        if (!ptr) return;
        auto it = live_.find((char*)ptr);
        if (it == live_.end()) throw Error(3, "device pool: free of unknown pointer");
        size_t ci = it->second.first, bytes = it->second.second;
        live_.erase(it);
        in_use_ -= bytes;
        auto& c = chunks_[ci];
        size_t off = (char*)ptr - c.base;
        auto ins = c.free_list.emplace(off, bytes).first;
        auto next = std::next(ins);
        if (next != c.free_list.end() && ins->first + ins->second == next->first) {
            ins->second += next->second;
            c.free_list.erase(next);
        }
        if (ins != c.free_list.begin()) {
            auto prev = std::prev(ins);
            if (prev->first + prev->second == ins->first) {
                prev->second += ins->second;
                c.free_list.erase(ins);
            }
        }
    }

    // Free every chunk that holds no live allocation.
    void trim() {
        std::vector<bool> used(chunks_.size(), false);
        for (auto& kv : live_) used[kv.second.first] = true;
        std::vector<Chunk> kept;
        std::vector<size_t> remap(chunks_.size(), (size_t)-1);
        for (size_t i = 0; i < chunks_.size(); i++) {
            if (used[i]) {
                remap[i] = kept.size();
                kept.push_back(std::move(chunks_[i]));
            } else {
                cudaFree(chunks_[i].base);
                capacity_ -= chunks_[i].size;
            }
        }
        chunks_ = std::move(kept);
        for (auto& kv : live_) kv.second.first = remap[kv.second.first];
    }

    void release_all() {
        for (auto& c : chunks_) cudaFree(c.base);
        chunks_.clear();
        live_.clear();
        in_use_ = 0;
        capacity_ = 0;
    }

    size_t capacity() const { return capacity_; }
    size_t in_use() const { return in_use_; }
    size_t peak() const { return peak_; }
    void reset_peak() { peak_ = in_use_; }

   private:
    struct Chunk {
        char* base = nullptr;
        size_t size = 0;
        std::map<size_t, size_t> free_list;  // offset -> length
    };
    static size_t round_up(size_t b) { return (b + kAlign - 1) / kAlign * kAlign; }
    void add_chunk(size_t bytes) {
        Chunk c;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw Error(2, "out of device memory: cudaMalloc(" + std::to_string(bytes) + " bytes) failed: " +
                               cudaGetErrorString(e));
        }
        c.base = (char*)p;
        c.size = bytes;
        c.free_list[0] = bytes;
        capacity_ += bytes;
        chunks_.push_back(std::move(c));
    }
    std::vector<Chunk> chunks_;
    std::map<char*, std::pair<size_t, size_t>> live_;  // ptr -> (chunk, bytes)
    size_t in_use_ = 0, peak_ = 0, capacity_ = 0;
};

// RAII handle on a pool allocation.
template <typename T>
class DevBuf {
   public:
    DevBuf() = default;
    DevBuf(DevicePool& pool, size_t count) : pool_(&pool), count_(count) { ptr_ = (T*)pool.alloc(count * sizeof(T)); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            reset();
            pool_ = o.pool_; ptr_ = o.ptr_; count_ = o.count_;
            o.ptr_ = nullptr; o.count_ = 0;
        }
        return *this;
    }
    ~DevBuf() { reset(); }
    void reset() {
        if (ptr_) pool_->free(ptr_);
        ptr_ = nullptr;
        count_ = 0;
    }
    T* release() { T* p = ptr_; ptr_ = nullptr; count_ = 0; return p; }
    T* get() const { return ptr_; }
    size_t size() const { return count_; }
    explicit operator bool() const { return ptr_ != nullptr; }

   private:
    DevicePool* pool_ = nullptr;
    T* ptr_ = nullptr;
    size_t count_ = 0;
};

}  // namespace sufr
