// ORACLE (test infrastructure) -- data model restating crates/loro-common + loro-internal op types.
// Never linked into the product (see oracle/README.md).
//   ID / ContainerID / ContainerType : crates/loro-common/src/lib.rs:28-47,114-180,293-347
//   LoroValue                         : crates/loro-common/src/value.rs
//   Op / InnerContent / Change        : crates/loro-internal/src/{op.rs,op/content.rs,change.rs}
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace lo {

typedef uint64_t PeerID;
typedef int32_t Counter;
typedef uint32_t Lamport;

struct ID {
    PeerID peer = 0;
    Counter counter = 0;
    bool operator==(const ID& o) const { return peer == o.peer && counter == o.counter; }
    bool operator!=(const ID& o) const { return !(*this == o); }
    bool operator<(const ID& o) const {
        return peer != o.peer ? peer < o.peer : counter < o.counter;
    }
    ID inc(int d) const { return ID{peer, counter + d}; }
};

// ContainerType::to_u8 (loro-common/src/lib.rs:293-347)
enum CType : uint8_t { CT_MAP = 0, CT_LIST = 1, CT_TEXT = 2, CT_TREE = 3, CT_MOVABLE = 4, CT_COUNTER = 5 };

struct ContainerID {
    bool root = true;
    std::string name;  // root
    PeerID peer = 0;   // normal
    Counter counter = 0;
    uint8_t type = 0;
    bool operator==(const ContainerID& o) const {
        return root == o.root && type == o.type &&
               (root ? name == o.name : (peer == o.peer && counter == o.counter));
    }
    bool operator<(const ContainerID& o) const {
        if (root != o.root) return root > o.root;
        if (type != o.type) return type < o.type;
        if (root) return name < o.name;
        if (peer != o.peer) return peer < o.peer;
        return counter < o.counter;
    }
};

struct Value {
    enum Kind : uint8_t { Null = 0, True = 1, False = 2, I64 = 3, F64 = 4, Str = 5, Binary = 6, List = 7, Map = 8, Container = 9 };
    Kind k = Null;
    int64_t i = 0;
    double f = 0;
    std::string s;                                   // Str / Binary
    std::vector<Value> list;                         // List
    std::vector<std::pair<std::string, Value>> map;  // Map (wire order preserved)
    ContainerID cid;                                 // Container
    static Value i64(int64_t v) { Value x; x.k = I64; x.i = v; return x; }
    static Value str(const std::string& v) { Value x; x.k = Str; x.s = v; return x; }
    static Value boolean(bool b) { Value x; x.k = b ? True : False; return x; }
    static Value f64(double d) { Value x; x.k = F64; x.f = d; return x; }
    static Value container(const ContainerID& c) { Value x; x.k = Container; x.cid = c; return x; }
    bool operator==(const Value& o) const {
        if (k != o.k) return false;
        switch (k) {
            case I64: return i == o.i;
            case F64: return std::memcmp(&f, &o.f, 8) == 0;
            case Str: case Binary: return s == o.s;
            case List: return list == o.list;
            case Map: return map == o.map;
            case Container: return cid == o.cid;
            default: return true;
        }
    }
};

// ValueKind tags of the `values` stream (encoding/value.rs:39-161)
enum VKind : uint8_t {
    VK_NULL = 0, VK_TRUE = 1, VK_FALSE = 2, VK_I64 = 3, VK_F64 = 4, VK_STR = 5, VK_BINARY = 6,
    VK_CONTAINER = 7, VK_DELETE_ONCE = 8, VK_DELETE_SEQ = 9, VK_DELTA_INT = 10, VK_LORO_VALUE = 11,
    VK_MARK_START = 12, VK_TREE_MOVE = 13, VK_LIST_MOVE = 14, VK_LIST_SET = 15, VK_RAW_TREE_MOVE = 16
};

enum OpKind : uint8_t {
    OP_LIST_INSERT,  // InnerListOp::Insert {slice,pos}
    OP_TEXT_INSERT,  // InnerListOp::InsertText
    OP_DELETE,       // InnerListOp::Delete(DeleteSpanWithId)
    OP_MAP_SET,      // MapSet{key, Some(v)}
    OP_MAP_DEL,      // MapSet{key, None}
    OP_TREE_CREATE, OP_TREE_MOVE, OP_TREE_DELETE,
    OP_STYLE_START, OP_STYLE_END, OP_LIST_MOVE, OP_LIST_SET,  // decoded, not merged
    OP_UNKNOWN
};

struct Op {
    Counter counter = 0;
    int cidx = -1;  // index into Doc::containers
    OpKind kind = OP_UNKNOWN;
    int32_t prop = 0;  // wire `prop` (pos / key idx / 0) -- for list/text ops == pos
    // list insert
    std::vector<Value> values;
    // text insert
    std::string text;
    uint32_t unicode_len = 0;
    // arena placement (models SharedArena adjacency, arena.rs:237-263,624-660; see doc.cpp)
    uint64_t arena_start = 0, arena_end = 0;   // values index (list) / byte offset (text)
    uint64_t arena_ustart = 0;                 // unicode offset (text)
    uint32_t arena_gen = 0;                    // AppendOnlyBytes buffer generation (text)
    // delete
    ID del_start;
    int64_t del_len = 0;  // signed
    // map
    std::string key;
    Value mapval;
    // tree
    ID target, parent;
    bool parent_null = true;
    std::string position;  // fractional index bytes
    // style / move / set (kept only so that blocks re-encode)
    uint32_t mark_len = 0; uint8_t mark_info = 0; std::string mark_key; Value mark_val;
    uint64_t mv_from = 0, mv_from_idx_peer = 0; uint64_t mv_lamport = 0; PeerID mv_peer = 0;
    uint8_t raw_vkind = 0;  // original value kind byte for unknown ops

    int atom_len() const {
        switch (kind) {
            case OP_LIST_INSERT: return (int)values.size();
            case OP_TEXT_INSERT: return (int)unicode_len;
            case OP_DELETE: return (int)(del_len < 0 ? -del_len : del_len);
            default: return 1;
        }
    }
    Counter ctr_end() const { return counter + atom_len(); }
    // DeleteSpan helpers (container/list/list_op.rs:290-376)
    int64_t del_pos() const { return prop; }
    int64_t del_start_pos() const { return del_len > 0 ? prop : prop + 1 + del_len; }
    bool del_bidirectional() const { return del_len == 1 || del_len == -1; }
    int64_t del_direction() const { return del_len > 0 ? 1 : -1; }
    int64_t del_next_pos() const { return del_len > 0 ? del_start_pos() : del_start_pos() - 1; }
    int64_t del_prev_pos() const { return del_len > 0 ? prop : prop + 1; }
    ID del_id_end() const { return del_start.inc((int)(del_len < 0 ? -del_len : del_len)); }
};

struct Change {
    ID id;
    Lamport lamport = 0;
    int64_t timestamp = 0;
    std::vector<ID> deps;
    bool has_msg = false;
    std::string msg;
    std::vector<Op> ops;
    int atom_len() const { return ops.empty() ? 0 : ops.back().ctr_end() - id.counter; }
    Counter ctr_end() const { return id.counter + atom_len(); }
    Lamport lamport_end() const { return lamport + (Lamport)atom_len(); }
    ID id_last() const { return ID{id.peer, ctr_end() - 1}; }
};

}  // namespace lo
