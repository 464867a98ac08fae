// lru.h — host-side per-endpoint LRU of block hashes.
//
// Mirrors upstream's podToLRU (capacity =
// lruCapacityPerServer, /root/reference/pkg/router/strategy.go:59,149) with
// hashicorp/golang-lru semantics: adding an existing key moves it to the front;
// adding a new key pushes it to the front and, if the set is full, evicts the
// oldest.  The LRU *order* is pointer-chasing and stays on the host; only the
// resulting membership changes (SET / CLEAR) are streamed to the GPU index.
//
// Flat arrays + an open-addressed key→node map with backward-shift deletion:
// ~40 bytes per entry, allocated on an endpoint's first use.
#pragma once
#include <cstdint>
#include <vector>

namespace fi {

class LruSet {
 public:
  explicit LruSet(uint32_t capacity = 0) : cap_(capacity) {}

  uint32_t size() const { return size_; }
  uint32_t capacity() const { return cap_; }

  // Touch `key`.  Returns true if it was newly inserted; *evicted/ *did_evict
  // report the key pushed out to make room.
  bool touch(uint64_t key, uint64_t* evicted, bool* did_evict) {
    *did_evict = false;
    if (cap_ == 0) return false;
    if (nodes_.empty()) init();
    uint32_t slot = find_slot(key);
    if (map_idx_[slot] != kNone) {  // hit: move to front
      move_front(map_idx_[slot]);
      return false;
    }
    uint32_t node;
    if (size_ == cap_) {  // evict the tail, reuse its node
      node = tail_;
      *evicted = nodes_[node].key;
      *did_evict = true;
      unlink(node);
      map_erase(nodes_[node].key);
      --size_;
      slot = find_slot(key);  // the erase may have shifted entries
    } else {
      node = size_;  // nodes are handed out densely until full
    }
    nodes_[node].key = key;
    link_front(node);
    map_key_[slot] = key;
    map_idx_[slot] = node;
    ++size_;
    return true;
  }

  bool contains(uint64_t key) const {
    if (nodes_.empty()) return false;
    return map_idx_[find_slot(key)] != kNone;
  }

 private:
  static constexpr uint32_t kNone = 0xFFFFFFFFu;
  struct Node {
    uint64_t key;
    uint32_t prev, next;
  };

  void init() {
    nodes_.resize(cap_);
    uint64_t m = 16;
    while (m < (uint64_t)cap_ * 2) m <<= 1;
    mask_ = (uint32_t)(m - 1);
    map_key_.assign(m, 0);
    map_idx_.assign(m, kNone);
    head_ = tail_ = kNone;
  }
  static inline uint64_t mix(uint64_t h) {
    h ^= h >> 31;
    h *= 0x9E3779B97F4A7C15ULL;
    return h ^ (h >> 29);
  }
  uint32_t find_slot(uint64_t key) const {
    uint32_t i = (uint32_t)mix(key) & mask_;
    while (map_idx_[i] != kNone && map_key_[i] != key) i = (i + 1) & mask_;
    return i;
  }
  void map_erase(uint64_t key) {
    uint32_t i = find_slot(key);
    if (map_idx_[i] == kNone) return;
    // backward-shift deletion keeps probe sequences intact without tombstones
    uint32_t j = i;
    for (;;) {
      j = (j + 1) & mask_;
      if (map_idx_[j] == kNone) break;
      uint32_t home = (uint32_t)mix(map_key_[j]) & mask_;
      // can entry j move into hole i?  yes iff home is not in (i, j] cyclically
      bool in_range = (i <= j) ? (home > i && home <= j) : (home > i || home <= j);
      if (!in_range) {
        map_key_[i] = map_key_[j];
        map_idx_[i] = map_idx_[j];
        i = j;
      }
    }
    map_idx_[i] = kNone;
  }
  void unlink(uint32_t n) {
    Node& x = nodes_[n];
    if (x.prev != kNone) nodes_[x.prev].next = x.next; else head_ = x.next;
    if (x.next != kNone) nodes_[x.next].prev = x.prev; else tail_ = x.prev;
  }
  void link_front(uint32_t n) {
    nodes_[n].prev = kNone;
    nodes_[n].next = head_;
    if (head_ != kNone) nodes_[head_].prev = n;
    head_ = n;
    if (tail_ == kNone) tail_ = n;
  }
  void move_front(uint32_t n) {
    if (head_ == n) return;
    unlink(n);
    link_front(n);
  }

  uint32_t cap_;
  uint32_t size_ = 0;
  uint32_t head_ = kNone, tail_ = kNone;
  uint32_t mask_ = 0;
  std::vector<Node> nodes_;
  std::vector<uint64_t> map_key_;
  std::vector<uint32_t> map_idx_;
};

}  // namespace fi
