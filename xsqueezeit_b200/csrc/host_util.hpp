// host_util.hpp -- host-side helpers shared by the CUDA context (xsi_b200.cu) and the
// container layer (xsi_container.cpp): dictionary order, tiny WAH codec for the per-block
// bool vectors, dictionary keys.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace xsi {

// Dictionary keys of a GT block (reference include/gt_block.hpp:36-60)
enum GtKey : uint32_t {
    KEY_BCF_LINES = 0, KEY_BINARY_LINES = 1, KEY_MAX_LINE_PLOIDY = 2, KEY_DEFAULT_PHASING = 3,
    KEY_WEIRDNESS_STRATEGY = 4,
    KEY_LINE_SORT = 0x10, KEY_LINE_SELECT = 0x11, KEY_LINE_HAPLOID = 0x12, KEY_LINE_MISSING = 0x16,
    KEY_LINE_NON_UNIFORM_PHASING = 0x17, KEY_LINE_END_OF_VECTORS = 0x18,
    KEY_MATRIX_WAH = 0x20, KEY_MATRIX_SPARSE = 0x21, KEY_MATRIX_MISSING = 0x26,
    KEY_MATRIX_NON_UNIFORM_PHASING = 0x27, KEY_MATRIX_END_OF_VECTORS = 0x28,
    KEY_MATRIX_MISSING_SPARSE = 0x36, KEY_MATRIX_END_OF_VECTORS_SPARSE = 0x38,
};
constexpr uint32_t VAL_UNDEFINED = 0xFFFFFFFFu;
constexpr uint32_t KEY_GT_ENTRY = 256;  // outer block dictionary, interfaces.hpp:167
constexpr uint32_t WS_SPARSE = 2;       // gt_block.hpp:69,417
constexpr uint32_t WS_WAH = 1;          // gt_block.hpp:68,174-176 (--wah-encode-missing)

// The reference serialises both per-block dictionaries by iterating a
// std::unordered_map<uint32_t,uint32_t> (interfaces.hpp:37-54; gt_block.hpp:461-510), so the
// on-disk order is libstdc++'s hash-table order for that insertion sequence.  To stay
// byte-exact independently of the C++ runtime this class replays that policy explicitly:
// identity hash, bucket counts 1 -> 13 -> next prime >= 2n, a node entering an empty bucket is
// pushed at the global list head, a node entering a non-empty bucket goes to that bucket's
// front; a rehash relinks the nodes in list order by the same rule.
class RefDictOrder {
public:
    void insert(uint32_t key) {
        for (const Node& n : nodes_) if (n.key == key) return;
        if (nodes_.size() + 1 > next_resize_) {
            size_t want = nodes_.size() + 1;
            if (next_resize_ == 0 && want < 11) want = 11;
            if (want >= nb_) rehash(next_bucket_count(std::max(want + 1, nb_ * 2)));
            else next_resize_ = nb_;
        }
        nodes_.push_back({key, -1});
        link((int)nodes_.size() - 1);
    }
    std::vector<uint32_t> order() const {
        std::vector<uint32_t> o;
        for (int p = head_; p >= 0; p = nodes_[p].next) o.push_back(nodes_[p].key);
        return o;
    }

private:
    struct Node { uint32_t key; int next; };
    static constexpr int EMPTY = -2, BEFORE_BEGIN = -1;
    std::vector<Node> nodes_;
    std::vector<int> bucket_ = std::vector<int>(1, EMPTY);  // node preceding the bucket's first node
    size_t nb_ = 1, next_resize_ = 0;
    int head_ = -1;

    size_t next_bucket_count(size_t n) {
        static const unsigned char fast[] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13};
        static const unsigned primes[] = {17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97,
                                          103, 109, 113, 127, 137, 139, 149, 157, 167, 179, 193, 199, 211, 227, 241,
                                          257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541, 577, 619, 661};
        size_t r = 0;
        if (n < 14) r = fast[n];
        else for (unsigned p : primes) if (p >= n) { r = p; break; }
        if (!r) r = n | 1;
        next_resize_ = r;
        return r;
    }
    void push_front_or_bucket(int p) {
        const size_t b = nodes_[p].key % nb_;
        if (bucket_[b] == EMPTY) {
            nodes_[p].next = head_;
            head_ = p;
            if (nodes_[p].next >= 0) bucket_[nodes_[nodes_[p].next].key % nb_] = p;
            bucket_[b] = BEFORE_BEGIN;
        } else {
            const int prev = bucket_[b];
            if (prev == BEFORE_BEGIN) { nodes_[p].next = head_; head_ = p; }
            else { nodes_[p].next = nodes_[prev].next; nodes_[prev].next = p; }
        }
    }
    void link(int p) { push_front_or_bucket(p); }
    void rehash(size_t nb) {
        std::vector<int> seq;
        for (int p = head_; p >= 0; p = nodes_[p].next) seq.push_back(p);
        nb_ = nb;
        bucket_.assign(nb, EMPTY);
        head_ = -1;
        size_t bbegin = 0;
        for (int p : seq) {
            const size_t b = nodes_[p].key % nb_;
            if (bucket_[b] == EMPTY) {
                nodes_[p].next = head_;
                head_ = p;
                bucket_[b] = BEFORE_BEGIN;
                if (nodes_[p].next >= 0) bucket_[bbegin] = p;
                bbegin = b;
            } else {
                const int prev = bucket_[b];
                if (prev == BEFORE_BEGIN) { nodes_[p].next = head_; head_ = p; }
                else { nodes_[p].next = nodes_[prev].next; nodes_[prev].next = p; }
            }
        }
    }
};

// WAH2-16 of a small bool vector (the per-block LINE_* vectors; reference wah.hpp:238-342)
inline void wah16_encode_bools(const std::vector<uint8_t>& bits, std::vector<uint8_t>& out) {
    auto put = [&](uint16_t w) { out.push_back((uint8_t)(w & 0xFF)); out.push_back((uint8_t)(w >> 8)); };
    uint16_t zeros = 0, ones = 0;
    const size_t n = bits.size(), groups = (n + 14) / 15;
    for (size_t g = 0; g < groups; ++g) {
        uint16_t w = 0;
        for (unsigned j = 0; j < 15; ++j) { const size_t i = g * 15 + j; if (i < n && bits[i]) w |= (uint16_t)(1u << j); }
        if (w == 0) {
            if (ones) { put(0xC000u | ones); ones = 0; }
            if (zeros == 0x3FFF) { put(0xBFFF); zeros = 0; }
            zeros++;
        } else if (w == 0x7FFF) {
            if (zeros) { put(0x8000u | zeros); zeros = 0; }
            if (ones == 0x3FFF) { put(0xFFFF); ones = 0; }
            ones++;
        } else {
            if (ones) { put(0xC000u | ones); ones = 0; }
            if (zeros) { put(0x8000u | zeros); zeros = 0; }
            put(w);
        }
    }
    if (zeros) put(0x8000u | zeros);
    if (ones) put(0xC000u | ones);
}

// Expands `size` bits the way wah2_extract does (whole words until >= size; wah.hpp:177-223),
// never reading past `end`.  Returns bytes consumed.
inline size_t wah16_decode_bools(const uint8_t* p, const uint8_t* end, size_t size, std::vector<uint8_t>& bits) {
    bits.assign(size, 0);
    size_t pos = 0;
    const uint8_t* q = p;
    while (pos < size && q + 2 <= end) {
        const uint16_t w = (uint16_t)(q[0] | (q[1] << 8));
        q += 2;
        if (w & 0x8000u) {
            const size_t len = (size_t)(w & 0x3FFFu) * 15;
            if (w & 0x4000u) for (size_t i = pos; i < pos + len && i < size; ++i) bits[i] = 1;
            pos += len;
        } else {
            for (unsigned j = 0; j < 15; ++j) if (pos + j < size) bits[pos + j] = (w >> j) & 1;
            pos += 15;
        }
    }
    return (size_t)(q - p);
}

inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t rd_u64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline void put_u32(std::vector<uint8_t>& b, uint32_t v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); b.insert(b.end(), p, p + 4); }

}  // namespace xsi
