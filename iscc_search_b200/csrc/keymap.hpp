// keymap.hpp - host-side exact key -> row-location map of a store (open addressing, linear probing,
// backward-shift deletion). Takes the place of iscc-usearch's per-shard bloom filters + key lookup
// (SURVEY.md 2.1): membership here is exact, which is what the reference relies on for
// `key in nphd_index` (/root/reference/iscc_search/indexes/usearch/index.py:560) and
// `composite_key in self._index` (/root/reference/iscc_search/indexes/simprint/usearch_core.py:135).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace isx {

struct Key128 {
    uint64_t hi, lo;  // 8-byte keys: (key, 0); 16-byte keys: big-endian halves
    bool operator==(const Key128& o) const { return hi == o.hi && lo == o.lo; }
};

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

class KeyMap {
   public:
    static constexpr uint64_t kEmpty = 0;  // stored location is loc + 1

    KeyMap() { rehash(1024); }

    size_t size() const { return size_; }

    void clear() {
        std::vector<Entry>().swap(tab_);
        size_ = 0;
        rehash(1024);
    }

    void reserve(size_t n) {
        size_t need = 1024;
        while (need * 5 < n * 8) need <<= 1;  // load factor <= 0.625
        if (need > tab_.size()) rehash(need);
    }

    inline size_t slot_of(const Key128& k) const { return (size_t)(mix64(k.hi + 0x9E3779B97F4A7C15ull) ^ mix64(k.lo)) & mask_; }
    inline void prefetch(const Key128& k) const { __builtin_prefetch(&tab_[slot_of(k)]); }

    // returns true and the location when present
    bool find(const Key128& k, uint64_t* loc) const {
        size_t i = slot_of(k);
        for (;;) {
            const Entry& e = tab_[i];
            if (e.loc1 == kEmpty) return false;
            if (e.key == k) { *loc = e.loc1 - 1; return true; }
            i = (i + 1) & mask_;
        }
    }

    // insert if absent; returns false when the key exists (nothing changed)
    bool insert(const Key128& k, uint64_t loc) {
        if ((size_ + 1) * 8 > tab_.size() * 5) rehash(tab_.size() * 2);
        size_t i = slot_of(k);
        for (;;) {
            Entry& e = tab_[i];
            if (e.loc1 == kEmpty) { e.key = k; e.loc1 = loc + 1; size_++; return true; }
            if (e.key == k) return false;
            i = (i + 1) & mask_;
        }
    }

    // overwrite the location of an existing key
    bool update(const Key128& k, uint64_t loc) {
        size_t i = slot_of(k);
        for (;;) {
            Entry& e = tab_[i];
            if (e.loc1 == kEmpty) return false;
            if (e.key == k) { e.loc1 = loc + 1; return true; }
            i = (i + 1) & mask_;
        }
    }

    bool erase(const Key128& k) {
        size_t i = slot_of(k);
        for (;;) {
            Entry& e = tab_[i];
            if (e.loc1 == kEmpty) return false;
            if (e.key == k) break;
            i = (i + 1) & mask_;
        }
        // backward shift: keep every remaining entry reachable from its home slot
        size_t hole = i;
        for (;;) {
            i = (i + 1) & mask_;
            Entry& e = tab_[i];
            if (e.loc1 == kEmpty) break;
            size_t home = slot_of(e.key);
            // can e move into the hole?  yes iff home is cyclically outside (hole, i]
            bool movable = (hole <= i) ? (home <= hole || home > i) : (home <= hole && home > i);
            if (movable) { tab_[hole] = e; hole = i; }
        }
        tab_[hole].loc1 = kEmpty;
        size_--;
        return true;
    }

   private:
    struct Entry { Key128 key; uint64_t loc1; };
    std::vector<Entry> tab_;
    size_t mask_ = 0, size_ = 0;

    void rehash(size_t cap) {
        std::vector<Entry> old;
        old.swap(tab_);
        tab_.assign(cap, Entry{{0, 0}, kEmpty});
        mask_ = cap - 1;
        size_t n = 0;
        for (const Entry& e : old) {
            if (e.loc1 == kEmpty) continue;
            size_t i = slot_of(e.key);
            while (tab_[i].loc1 != kEmpty) i = (i + 1) & mask_;
            tab_[i] = e;
            n++;
        }
        size_ = n;
    }
};

}  // namespace isx
