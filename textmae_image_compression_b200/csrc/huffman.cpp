// Side-information coder of the reference's eval CLI: utils/huffman.py HuffmanCoding, used by testing.py:73-76
// to code `ids_restore` and to add len(bitstring) / pixels to the reported bpp (testing.py:89).  Host-only, not on the
// device path (north_star keeps bitstream emission outside the hot path); it lives behind the C ABI so a host that links
// libtmae_b200.so gets the reference's exact bit string without Python.
//
// Bit-exactness with the reference needs its tie-breaking, which is CPython's heapq on nodes compared by frequency only
// (utils/huffman.py:28-38): the binary-heap sift routines below restate heapq.heappush / heappop (Lib/heapq.py: _siftdown,
// _siftup - "bubble the smaller child up to a leaf, then sift down", comparisons with `<` only), the frequency table keeps
// first-appearance order (a dict, :55-62), the tree is walked left = "0", right = "1" (:76-93).
#include <stdint.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tmae.h"

struct tmae_huffman {
    struct Node { int64_t value; int64_t freq; int left, right; bool leaf; };
    std::vector<Node> nodes;
    std::vector<int> heap;
    std::unordered_map<int64_t, std::string> codes;          // value -> code (utils/huffman.py `codes`)
    std::unordered_map<std::string, int64_t> reverse;        // code -> value (`reverse_mapping`)
    std::vector<int64_t> order;                              // distinct values in code-assignment (pre-order) order = key order of `codes`
    std::string bits;                                        // '0' / '1' characters: the reference's encoded_text
    std::string err;

    bool lt(int a, int b) const { return nodes[a].freq < nodes[b].freq; }      // Node.__lt__
    void siftdown(int startpos, int pos) {                   // heapq._siftdown
        const int newitem = heap[pos];
        while (pos > startpos) {
            const int parentpos = (pos - 1) >> 1;
            const int parent = heap[parentpos];
            if (lt(newitem, parent)) { heap[pos] = parent; pos = parentpos; continue; }
            break;
        }
        heap[pos] = newitem;
    }
    void siftup(int pos) {                                   // heapq._siftup
        const int endpos = (int)heap.size(), startpos = pos;
        const int newitem = heap[pos];
        int childpos = 2 * pos + 1;
        while (childpos < endpos) {
            const int rightpos = childpos + 1;
            if (rightpos < endpos && !lt(heap[childpos], heap[rightpos])) childpos = rightpos;
            heap[pos] = heap[childpos];
            pos = childpos;
            childpos = 2 * pos + 1;
        }
        heap[pos] = newitem;
        siftdown(startpos, pos);
    }
    void push(int n) { heap.push_back(n); siftdown(0, (int)heap.size() - 1); }
    int pop() {
        const int last = heap.back();
        heap.pop_back();
        if (heap.empty()) return last;
        const int ret = heap[0];
        heap[0] = last;
        siftup(0);
        return ret;
    }
    void build(const int64_t* v, int64_t n) {
        nodes.clear(); heap.clear(); codes.clear(); reverse.clear(); order.clear(); bits.clear();
        std::unordered_map<int64_t, int> slot;
        for (int64_t i = 0; i < n; ++i) {                    // build_heap: frequencies in first-appearance order (:55-58)
            auto it = slot.find(v[i]);
            if (it == slot.end()) { slot.emplace(v[i], (int)nodes.size()); nodes.push_back({v[i], 1, -1, -1, true}); }
            else nodes[it->second].freq++;
        }
        const int leaves = (int)nodes.size();
        for (int k = 0; k < leaves; ++k) push(k);            // :60-62
        while (heap.size() > 1) {                            // build_tree (:68-76)
            const int a = pop(), b = pop();
            nodes.push_back({0, nodes[a].freq + nodes[b].freq, a, b, false});
            push((int)nodes.size() - 1);
        }
        if (heap.empty()) return;
        // build_codes (:95-102): pre-order walk, left "0" / right "1"; explicit stack instead of recursion
        std::vector<std::pair<int, std::string>> st;
        st.emplace_back(heap[0], std::string());
        while (!st.empty()) {
            auto cur = std::move(st.back());
            st.pop_back();
            const Node& nd = nodes[cur.first];
            if (nd.leaf) { codes[nd.value] = cur.second; reverse[cur.second] = nd.value; order.push_back(nd.value); continue; }
            st.emplace_back(nd.right, cur.second + "1");
            st.emplace_back(nd.left, cur.second + "0");
        }
    }
};

namespace {
thread_local std::string g_huff_error;
int hfail(tmae_huffman* h, int code, const char* msg) {
    if (h) h->err = msg; else g_huff_error = msg;
    return code;
}
}  // namespace

extern "C" {

int tmae_huffman_create(tmae_huffman** out) {
    if (!out) return hfail(nullptr, TMAE_EINVAL, "null output pointer");
    *out = new (std::nothrow) tmae_huffman();
    return *out ? TMAE_OK : hfail(nullptr, TMAE_ENOMEM, "out of memory");
}
void tmae_huffman_destroy(tmae_huffman* h) { delete h; }
const char* tmae_huffman_last_error(const tmae_huffman* h) { return h ? h->err.c_str() : g_huff_error.c_str(); }

int tmae_huffman_compress(tmae_huffman* h, const int64_t* h_values, int64_t n, int64_t* n_bits) {
    if (!h || n < 0 || (n > 0 && !h_values)) return hfail(h, TMAE_EINVAL, "invalid argument");
    try {
        h->build(h_values, n);
        size_t total = 0;
        for (int64_t i = 0; i < n; ++i) total += h->codes[h_values[i]].size();
        h->bits.reserve(total);
        for (int64_t i = 0; i < n; ++i) h->bits += h->codes[h_values[i]];          // encode (:104-118)
    } catch (const std::bad_alloc&) { return hfail(h, TMAE_ENOMEM, "out of memory"); }
    if (n_bits) *n_bits = (int64_t)h->bits.size();
    return TMAE_OK;
}

int tmae_huffman_bits(const tmae_huffman* h, uint8_t* h_out, int64_t capacity, int as_chars) {
    if (!h || (!h_out && capacity > 0)) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "invalid argument");
    const int64_t nb = (int64_t)h->bits.size();
    const int64_t need = as_chars ? nb : (nb + 7) / 8;
    if (capacity < need) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "output buffer too small");
    if (as_chars) { memcpy(h_out, h->bits.data(), (size_t)nb); return TMAE_OK; }
    memset(h_out, 0, (size_t)need);
    for (int64_t i = 0; i < nb; ++i) if (h->bits[(size_t)i] == '1') h_out[i >> 3] |= (uint8_t)(0x80u >> (i & 7));
    return TMAE_OK;
}

int tmae_huffman_code(const tmae_huffman* h, int64_t value, char* out, int capacity) {
    if (!h || !out) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "invalid argument");
    auto it = h->codes.find(value);
    if (it == h->codes.end()) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "value has no code");
    if ((int)it->second.size() + 1 > capacity) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "output buffer too small");
    memcpy(out, it->second.c_str(), it->second.size() + 1);
    return TMAE_OK;
}

int tmae_huffman_num_symbols(const tmae_huffman* h, int64_t* h_values, int64_t capacity) {
    if (!h) return -1;
    if (h_values) for (int64_t i = 0; i < capacity && i < (int64_t)h->order.size(); ++i) h_values[i] = h->order[(size_t)i];
    return (int)h->order.size();
}

int tmae_huffman_decompress(const tmae_huffman* h, const uint8_t* h_bits, int64_t n_bits, int as_chars, int64_t* h_values,
                            int64_t capacity, int64_t* n_values) {
    if (!h || n_bits < 0 || (n_bits > 0 && !h_bits) || !n_values) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "invalid argument");
    // decode (:120-139): grow the current code bit by bit, emit on a match
    std::string cur;
    int64_t count = 0;
    for (int64_t i = 0; i < n_bits; ++i) {
        const bool one = as_chars ? (h_bits[i] == '1') : ((h_bits[i >> 3] >> (7 - (i & 7))) & 1u);
        cur.push_back(one ? '1' : '0');
        auto it = h->reverse.find(cur);
        if (it != h->reverse.end()) {
            if (count < capacity && h_values) h_values[count] = it->second;
            ++count;
            cur.clear();
        }
    }
    *n_values = count;
    if (count > capacity) return hfail(const_cast<tmae_huffman*>(h), TMAE_EINVAL, "output buffer too small");
    return TMAE_OK;
}

}  // extern "C"
