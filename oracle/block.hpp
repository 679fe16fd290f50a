// ORACLE (test infrastructure) -- change-block + blob codec, restating
//   crates/loro-internal/src/oplog/change_store/block_encode.rs:95-119 (EncodedBlock),
//     :137-278 (encode_block), :527-659 (decode_block)
//   .../block_meta_encode.rs:13-88 (encode_changes), :90-179 (decode_changes_header)
//   crates/loro-internal/src/encoding/value.rs:343-458,603-990,992-1135 (values stream)
//   .../encoding/outdated_encode_reordered.rs:104-211 (get_op_prop/encode_op), :215-423 (decode_op)
//   .../encoding/arena.rs:39-147 (ContainerArena), :159-233 (PositionArena)
//   crates/loro-internal/src/encoding.rs:278-330,397-416 (header + checksum)
//   .../encoding/fast_snapshot.rs:257-288 (FastUpdates framing)
// Never linked into the product.
#pragma once
#include <algorithm>
#include <set>
#include "codec.hpp"
#include "model.hpp"

namespace lo {

// What decode/encode need from the owning document (SharedArena in the reference).
struct ArenaCtx {
    virtual int register_container(const ContainerID& c) = 0;
    virtual const ContainerID& container_id(int cidx) const = 0;
    virtual void alloc_values(Op& op) = 0;  // arena.rs:260-263
    virtual void alloc_str(Op& op) = 0;     // arena.rs:237-240
    virtual ~ArenaCtx() {}
};

template <class T>
struct Register {  // encoding/value_register.rs: first-use order
    std::vector<T> vec;
    std::map<T, size_t> idx;
    size_t reg(const T& v) {
        auto it = idx.find(v);
        if (it != idx.end()) return it->second;
        size_t i = vec.size();
        vec.push_back(v);
        idx[v] = i;
        return i;
    }
};

static const PeerID DELETED_TREE_ROOT_PEER = UINT64_MAX;   // loro-common/src/lib.rs:631
static const Counter DELETED_TREE_ROOT_CTR = INT32_MAX;

inline size_t utf8_chars(const std::string& s) {
    size_t n = 0;
    for (unsigned char c : s) n += ((c & 0xC0) != 0x80);
    return n;
}

// ------------------------------------------------------------------ LoroValue in `values`
inline Value read_loro_value_content(Reader& r, uint8_t kind, const std::vector<std::string>& keys,
                                     ID id, bool top_list);
inline Value read_loro_value(Reader& r, const std::vector<std::string>& keys, ID id) {
    uint8_t kind = r.u8();
    return read_loro_value_content(r, kind, keys, id, false);
}
inline Value read_loro_value_content(Reader& r, uint8_t kind, const std::vector<std::string>& keys,
                                     ID id, bool top_list) {
    Value v;
    switch (kind) {
        case 0: v.k = Value::Null; break;
        case 1: v.k = Value::True; break;
        case 2: v.k = Value::False; break;
        case 3: v.k = Value::I64; v.i = r.sleb(); break;
        case 4: {
            v.k = Value::F64;
            const uint8_t* b = r.take(8);
            uint64_t bits = 0;
            for (int i = 0; i < 8; i++) bits = (bits << 8) | b[i];  // big endian (value.rs:874-883)
            std::memcpy(&v.f, &bits, 8);
            break;
        }
        case 5: case 6: {
            v.k = kind == 5 ? Value::Str : Value::Binary;
            uint64_t n = r.uleb();
            const uint8_t* b = r.take(n);
            v.s.assign((const char*)b, n);
            break;
        }
        case 7: {
            v.k = Value::List;
            uint64_t n = r.uleb();
            if (n > (1u << 28)) throw DecodeError("collection too large");
            for (uint64_t i = 0; i < n; i++)
                v.list.push_back(read_loro_value(r, keys, top_list ? id.inc((int)i) : id));
            break;
        }
        case 8: {
            v.k = Value::Map;
            uint64_t n = r.uleb();
            if (n > (1u << 28)) throw DecodeError("collection too large");
            for (uint64_t i = 0; i < n; i++) {
                uint64_t ki = r.uleb();
                if (ki >= keys.size()) throw DecodeError("bad key idx");
                Value x = read_loro_value(r, keys, id);
                v.map.push_back({keys[ki], x});
            }
            break;
        }
        case 9: {
            v.k = Value::Container;
            uint8_t t = r.u8();
            v.cid.root = false;
            v.cid.peer = id.peer;
            v.cid.counter = id.counter;
            v.cid.type = t;
            break;
        }
        default: throw DecodeError("bad LoroValueKind");
    }
    return v;
}
inline void write_loro_value(Writer& w, const Value& v, Register<std::string>& keys) {
    w.u8((uint8_t)v.k);
    switch (v.k) {
        case Value::I64: w.sleb(v.i); break;
        case Value::F64: {
            uint64_t bits;
            std::memcpy(&bits, &v.f, 8);
            for (int i = 7; i >= 0; i--) w.u8((uint8_t)(bits >> (8 * i)));
            break;
        }
        case Value::Str: case Value::Binary:
            w.uleb(v.s.size());
            w.bytes(v.s);
            break;
        case Value::List:
            w.uleb(v.list.size());
            for (auto& x : v.list) write_loro_value(w, x, keys);
            break;
        case Value::Map:
            w.uleb(v.map.size());
            for (auto& kv : v.map) {
                w.uleb(keys.reg(kv.first));
                write_loro_value(w, kv.second, keys);
            }
            break;
        case Value::Container: w.u8(v.cid.type); break;
        default: break;
    }
}

// ------------------------------------------------------------------ block decode
struct BlockMeta {  // what tests inspect (ChangesBlockHeader, block_encode.rs:370-388)
    uint32_t counter_start = 0, counter_len = 0, lamport_start = 0, lamport_len = 0, n_changes = 0;
    std::vector<PeerID> peers;
    std::vector<std::string> keys;
    std::vector<ContainerID> cids;
    size_t sec_len[8] = {0};  // header, change_meta, cids, keys, positions, ops, delete_start_ids, values
};

inline std::vector<Change> decode_block(const uint8_t* bytes, size_t n, ArenaCtx& ctx,
                                        BlockMeta* meta_out = nullptr) {
    Reader r(bytes, n);
    BlockMeta m;
    m.counter_start = (uint32_t)r.varint();
    m.counter_len = (uint32_t)r.varint();
    m.lamport_start = (uint32_t)r.varint();
    m.lamport_len = (uint32_t)r.varint();
    m.n_changes = (uint32_t)r.varint();
    const uint8_t* sec[8];
    for (int i = 0; i < 8; i++) {
        m.sec_len[i] = (size_t)r.varint();
        sec[i] = r.take(m.sec_len[i]);
    }
    if (!r.empty()) throw DecodeError("block: trailing bytes");
    size_t N = m.n_changes;
    if (N == 0) throw DecodeError("block: no changes");

    // ---- header (block_meta_encode.rs:90-179)
    Reader h(sec[0], m.sec_len[0]);
    uint64_t peer_num = h.uleb();
    for (uint64_t i = 0; i < peer_num; i++) {
        const uint8_t* b = h.take(8);
        PeerID p = 0;
        for (int k = 7; k >= 0; k--) p = (p << 8) | b[k];
        m.peers.push_back(p);
    }
    if (m.peers.empty()) throw DecodeError("block: no peers");
    std::vector<Counter> lengths;
    int64_t sum = 0;
    for (size_t i = 0; i + 1 < N; i++) {
        lengths.push_back((Counter)h.uleb());
        sum += lengths.back();
    }
    lengths.push_back((Counter)((int64_t)m.counter_len - sum));
    std::vector<bool> dep_self = bool_rle_take_n(h, N);
    std::vector<uint64_t> deps_len;
    any_rle_take_n<uint64_t>(h, N, deps_len, RdVar());
    size_t other_dep_num = 0;
    for (auto x : deps_len) other_dep_num += x;
    std::vector<uint64_t> dep_peers;
    any_rle_take_n<uint64_t>(h, other_dep_num, dep_peers, RdVar());
    std::vector<int64_t> dep_counters = dod_take_n(h, other_dep_num);
    std::vector<int64_t> lamports = dod_take_n(h, N - 1);
    lamports.push_back((int64_t)m.lamport_start + m.lamport_len - (uint32_t)lengths.back());
    if (!h.empty()) throw DecodeError("header: trailing bytes");

    // ---- change meta (block_encode.rs:553-556)
    Reader cm(sec[1], m.sec_len[1]);
    std::vector<int64_t> timestamps = dod_take_n(cm, N);
    std::vector<uint64_t> msg_lens;
    any_rle_take_n<uint64_t>(cm, N, msg_lens, RdVar());

    // ---- keys (block_encode.rs:290-300)
    {
        Reader k(sec[3], m.sec_len[3]);
        while (!k.empty()) {
            uint64_t len = k.uleb();
            const uint8_t* b = k.take(len);
            m.keys.push_back(std::string((const char*)b, len));
        }
    }
    // ---- cids (arena.rs:94-101; row-wise postcard, each row prefixed by field count 4)
    {
        Reader c(sec[2], m.sec_len[2]);
        uint64_t nc = c.varint();
        for (uint64_t i = 0; i < nc; i++) {
            if (c.varint() != 4) throw DecodeError("cids: field count");
            ContainerID id;
            uint8_t is_root = c.u8();
            id.type = c.u8();
            uint64_t peer_idx = c.varint();
            int64_t koc = c.zigzag();
            id.root = is_root != 0;
            if (id.root) {
                if (koc < 0 || (size_t)koc >= m.keys.size()) throw DecodeError("cids: key idx");
                id.name = m.keys[(size_t)koc];
            } else {
                if (peer_idx >= m.peers.size()) throw DecodeError("cids: peer idx");
                id.peer = m.peers[peer_idx];
                id.counter = (Counter)koc;
            }
            m.cids.push_back(id);
        }
        if (!c.empty()) throw DecodeError("cids: trailing");
    }
    // ---- positions (arena.rs:159-233)
    std::vector<std::string> positions;
    if (m.sec_len[4]) {
        auto cols = columnar_take_wrapped(sec[4], m.sec_len[4], 2);
        Reader c0(cols[0].first, cols[0].second);
        std::vector<uint64_t> common;
        any_rle_decode_all<uint64_t>(c0, common, RdVar());
        Reader c1(cols[1].first, cols[1].second);
        uint64_t np = c1.varint();
        if (np != common.size()) throw DecodeError("positions: count mismatch");
        std::string last;
        for (uint64_t i = 0; i < np; i++) {
            uint64_t len = c1.varint();
            const uint8_t* b = c1.take(len);
            if (common[i] > last.size()) throw DecodeError("positions: prefix");
            std::string p = last.substr(0, common[i]) + std::string((const char*)b, len);
            positions.push_back(p);
            last = p;
        }
    }
    // ---- ops columns (block_encode.rs:412-429,574)
    auto ocols = columnar_take_wrapped(sec[5], m.sec_len[5], 4);
    Reader oc0(ocols[0].first, ocols[0].second), oc1(ocols[1].first, ocols[1].second),
        oc2(ocols[2].first, ocols[2].second), oc3(ocols[3].first, ocols[3].second);
    std::vector<int64_t> col_cidx = delta_rle_decode_all(oc0);
    std::vector<int64_t> col_prop = delta_rle_decode_all(oc1);
    std::vector<uint64_t> col_vt, col_len;
    any_rle_decode_all<uint64_t>(oc2, col_vt, RdU8());
    any_rle_decode_all<uint64_t>(oc3, col_len, RdVar());
    size_t n_ops = col_cidx.size();
    if (col_prop.size() != n_ops || col_vt.size() != n_ops || col_len.size() != n_ops)
        throw DecodeError("ops: column length mismatch");
    // ---- delete start ids (outdated_encode_reordered.rs:427-436)
    std::vector<int64_t> d_peer, d_ctr, d_len;
    if (m.sec_len[6]) {
        auto dcols = columnar_take_wrapped(sec[6], m.sec_len[6], 3);
        Reader a(dcols[0].first, dcols[0].second), b(dcols[1].first, dcols[1].second),
            c(dcols[2].first, dcols[2].second);
        d_peer = delta_rle_decode_all(a);
        d_ctr = delta_rle_decode_all(b);
        d_len = delta_rle_decode_all(c);
    }
    size_t del_i = 0;

    // ---- assemble changes (block_encode.rs:585-657)
    std::vector<Change> changes(N);
    {
        Counter c = (Counter)m.counter_start;
        size_t dp = 0, msg_off = 0;
        const uint8_t* msgs = cm.p;
        size_t msgs_len = cm.remaining();
        for (size_t i = 0; i < N; i++) {
            Change& ch = changes[i];
            ch.id = ID{m.peers[0], c};
            ch.lamport = (Lamport)lamports[i];
            ch.timestamp = timestamps[i];
            if (dep_self[i]) ch.deps.push_back(ID{m.peers[0], c - 1});
            for (uint64_t k = 0; k < deps_len[i]; k++, dp++) {
                if (dep_peers[dp] >= m.peers.size()) throw DecodeError("deps: peer idx");
                ch.deps.push_back(ID{m.peers[dep_peers[dp]], (Counter)dep_counters[dp]});
            }
            if (msg_lens[i]) {
                if (msg_off + msg_lens[i] > msgs_len) throw DecodeError("msg: eof");
                ch.has_msg = true;
                ch.msg.assign((const char*)msgs + msg_off, msg_lens[i]);
                msg_off += msg_lens[i];
            }
            c += lengths[i];
        }
    }
    std::vector<Counter> counters;  // N+1 boundaries
    {
        Counter c = (Counter)m.counter_start;
        for (size_t i = 0; i < N; i++) {
            counters.push_back(c);
            c += lengths[i];
        }
        counters.push_back((Counter)(m.counter_start + m.counter_len));
    }

    Reader vr(sec[7], m.sec_len[7]);
    Counter counter = (Counter)m.counter_start;
    size_t change_index = 0;
    PeerID peer = m.peers[0];
    std::vector<int> cidx_map(m.cids.size(), -1);
    for (size_t i = 0; i < n_ops; i++) {
        if (col_cidx[i] < 0 || (size_t)col_cidx[i] >= m.cids.size()) throw DecodeError("op: cid idx");
        const ContainerID& cid = m.cids[(size_t)col_cidx[i]];
        int32_t prop = (int32_t)col_prop[i];
        uint8_t vt = (uint8_t)col_vt[i];
        ID op_id{peer, counter};
        Op op;
        op.counter = counter;
        op.prop = prop;
        op.raw_vkind = vt;
        auto take_del = [&](Op& o) {
            if (del_i >= d_peer.size()) throw DecodeError("delete ids exhausted");
            if (d_peer[del_i] < 0 || (size_t)d_peer[del_i] >= m.peers.size())
                throw DecodeError("delete: peer idx");
            o.kind = OP_DELETE;
            o.del_start = ID{m.peers[(size_t)d_peer[del_i]], (Counter)d_ctr[del_i]};
            o.del_len = d_len[del_i];
            if (o.del_len == 0) throw DecodeError("delete: zero len");
            del_i++;
        };
        // Value::decode (value.rs:343-391) fused with decode_op (outdated_encode_reordered.rs:215-423)
        switch (vt) {
            case VK_STR: {
                uint64_t len = vr.uleb();
                const uint8_t* b = vr.take(len);
                op.text.assign((const char*)b, len);
                op.unicode_len = (uint32_t)utf8_chars(op.text);
                if (cid.type != CT_TEXT) throw DecodeError("Str on non-text");
                op.kind = OP_TEXT_INSERT;
                break;
            }
            case VK_DELETE_SEQ: take_del(op); break;
            case VK_DELETE_ONCE:
                op.kind = OP_MAP_DEL;
                if (prop < 0 || (size_t)prop >= m.keys.size()) throw DecodeError("map key idx");
                op.key = m.keys[(size_t)prop];
                break;
            case VK_LORO_VALUE: {
                uint8_t kind = vr.u8();
                Value v = read_loro_value_content(vr, kind, m.keys, op_id,
                                                  cid.type == CT_LIST || cid.type == CT_MOVABLE);
                if (cid.type == CT_MAP) {
                    op.kind = OP_MAP_SET;
                    if (prop < 0 || (size_t)prop >= m.keys.size()) throw DecodeError("map key idx");
                    op.key = m.keys[(size_t)prop];
                    op.mapval = v;
                } else if (cid.type == CT_LIST || cid.type == CT_MOVABLE) {
                    if (v.k != Value::List) throw DecodeError("list insert: not a list");
                    op.kind = OP_LIST_INSERT;
                    op.values = v.list;
                } else
                    throw DecodeError("LoroValue on bad container");
                break;
            }
            case VK_NULL: op.kind = OP_STYLE_END; break;
            case VK_MARK_START: {
                op.kind = OP_STYLE_START;
                op.mark_info = vr.u8();
                op.mark_len = (uint32_t)vr.uleb();
                uint64_t ki = vr.uleb();
                if (ki >= m.keys.size()) throw DecodeError("mark key");
                op.mark_key = m.keys[ki];
                op.mark_val = read_loro_value(vr, m.keys, op_id);
                break;
            }
            case VK_RAW_TREE_MOVE: {
                uint64_t sp = vr.uleb(), sc = vr.uleb(), pi = vr.uleb();
                uint8_t pn = vr.u8();
                uint64_t pp = 0, pc = 0;
                if (!pn) { pp = vr.uleb(); pc = vr.uleb(); }
                if (sp >= m.peers.size() || (!pn && pp >= m.peers.size())) throw DecodeError("tree peer idx");
                op.target = ID{m.peers[sp], (Counter)sc};
                op.parent_null = pn != 0;
                if (!pn) op.parent = ID{m.peers[pp], (Counter)pc};
                if (!pn && op.parent.peer == DELETED_TREE_ROOT_PEER && op.parent.counter == DELETED_TREE_ROOT_CTR) {
                    op.kind = OP_TREE_DELETE;
                } else {
                    if (pi >= positions.size()) throw DecodeError("tree position idx");
                    op.position = positions[pi];
                    op.kind = (op.target == op_id) ? OP_TREE_CREATE : OP_TREE_MOVE;
                }
                break;
            }
            case VK_LIST_MOVE: {
                op.kind = OP_LIST_MOVE;
                op.mv_from = vr.uleb();
                uint64_t fi = vr.uleb();
                op.mv_lamport = vr.uleb();
                if (fi >= m.peers.size()) throw DecodeError("move peer idx");
                op.mv_peer = m.peers[fi];
                break;
            }
            case VK_LIST_SET: {
                op.kind = OP_LIST_SET;
                uint64_t pi = vr.uleb();
                op.mv_lamport = vr.uleb();
                if (pi >= m.peers.size()) throw DecodeError("set peer idx");
                op.mv_peer = m.peers[pi];
                op.mark_val = read_loro_value(vr, m.keys, op_id);
                break;
            }
            case VK_I64: op.kind = OP_UNKNOWN; op.mapval = Value::i64(vr.sleb()); break;  // counter
            case VK_F64: {
                op.kind = OP_UNKNOWN;
                const uint8_t* b = vr.take(8);
                uint64_t bits = 0;
                for (int k = 0; k < 8; k++) bits = (bits << 8) | b[k];
                double d; std::memcpy(&d, &bits, 8);
                op.mapval = Value::f64(d);
                break;
            }
            default: throw DecodeError("unsupported value kind " + std::to_string(vt));
        }
        if (cidx_map[(size_t)col_cidx[i]] < 0) cidx_map[(size_t)col_cidx[i]] = ctx.register_container(cid);
        op.cidx = cidx_map[(size_t)col_cidx[i]];
        if (op.kind == OP_TEXT_INSERT) ctx.alloc_str(op);
        if (op.kind == OP_LIST_INSERT) ctx.alloc_values(op);
        if ((uint64_t)op.atom_len() != col_len[i]) throw DecodeError("op: len column mismatch");
        if (change_index >= N) throw DecodeError("op: beyond last change");
        // changes[change_index].ops.push(op) is an RleVec push (merges mergable neighbours);
        // merging is done by the caller (doc.hpp push_op) so that this file stays a pure codec.
        changes[change_index].ops.push_back(op);
        counter += (Counter)col_len[i];
        if (counter >= counters[change_index + 1]) change_index++;
    }
    if (!vr.empty()) throw DecodeError("values: trailing bytes");
    if (meta_out) *meta_out = m;
    return changes;
}

// ------------------------------------------------------------------ block encode
struct EncodedSections {
    std::vector<uint8_t> sec[8];
};

inline std::vector<uint8_t> encode_block(const std::vector<Change>& block, const ArenaCtx& ctx,
                                         EncodedSections* out_secs = nullptr) {
    if (block.empty()) throw std::runtime_error("empty block");
    Register<PeerID> peers;
    Register<std::string> keys;
    Register<ContainerID> cids;
    Register<std::string> positions;
    PeerID peer = block[0].id.peer;
    peers.reg(peer);
    {  // block_encode.rs:156-178 : positions pre-registered in sorted order
        std::set<std::string> ps;
        for (auto& c : block)
            for (auto& op : c.ops)
                if (op.kind == OP_TREE_CREATE || op.kind == OP_TREE_MOVE) ps.insert(op.position);
        for (auto& p : ps) positions.reg(p);
    }
    std::vector<int64_t> c_cidx, c_prop, d_peer, d_ctr, d_len;
    std::vector<uint64_t> c_vt, c_len;
    Writer vw;
    for (auto& c : block) {
        for (auto& op : c.ops) {
            size_t cidx = cids.reg(ctx.container_id(op.cidx));
            int32_t prop = op.prop;
            if (op.kind == OP_MAP_SET || op.kind == OP_MAP_DEL) prop = (int32_t)keys.reg(op.key);
            uint8_t vk = 0;
            switch (op.kind) {
                case OP_LIST_INSERT: {
                    vk = VK_LORO_VALUE;
                    vw.u8(7);
                    vw.uleb(op.values.size());
                    for (auto& v : op.values) write_loro_value(vw, v, keys);
                    break;
                }
                case OP_TEXT_INSERT:
                    vk = VK_STR;
                    vw.uleb(op.text.size());
                    vw.bytes(op.text);
                    break;
                case OP_DELETE:
                    vk = VK_DELETE_SEQ;
                    d_peer.push_back((int64_t)peers.reg(op.del_start.peer));
                    d_ctr.push_back(op.del_start.counter);
                    d_len.push_back(op.del_len);
                    break;
                case OP_MAP_SET:
                    vk = VK_LORO_VALUE;
                    write_loro_value(vw, op.mapval, keys);
                    break;
                case OP_MAP_DEL: vk = VK_DELETE_ONCE; break;
                case OP_TREE_CREATE: case OP_TREE_MOVE: {
                    vk = VK_RAW_TREE_MOVE;
                    vw.uleb(peers.reg(op.target.peer));
                    vw.uleb((uint64_t)(uint32_t)op.target.counter);
                    uint64_t ppi = op.parent_null ? 0 : peers.reg(op.parent.peer);
                    vw.uleb(positions.reg(op.position));
                    vw.u8(op.parent_null ? 1 : 0);
                    if (!op.parent_null) {
                        vw.uleb(ppi);
                        vw.uleb((uint64_t)(uint32_t)op.parent.counter);
                    }
                    break;
                }
                case OP_TREE_DELETE: {
                    vk = VK_RAW_TREE_MOVE;
                    vw.uleb(peers.reg(op.target.peer));
                    vw.uleb((uint64_t)(uint32_t)op.target.counter);
                    uint64_t ppi = peers.reg(DELETED_TREE_ROOT_PEER);
                    vw.uleb(0);
                    vw.u8(0);
                    vw.uleb(ppi);
                    vw.uleb((uint64_t)(uint32_t)DELETED_TREE_ROOT_CTR);
                    break;
                }
                case OP_STYLE_START: {
                    vk = VK_MARK_START;
                    size_t ki = keys.reg(op.mark_key);  // write_mark registers the key first
                    vw.u8(op.mark_info);
                    vw.uleb(op.mark_len);
                    vw.uleb(ki);
                    write_loro_value(vw, op.mark_val, keys);
                    break;
                }
                case OP_STYLE_END: vk = VK_NULL; break;
                case OP_LIST_MOVE:
                    vk = VK_LIST_MOVE;
                    vw.uleb(op.mv_from);
                    vw.uleb(peers.reg(op.mv_peer));
                    vw.uleb(op.mv_lamport);
                    break;
                case OP_LIST_SET:
                    vk = VK_LIST_SET;
                    vw.uleb(peers.reg(op.mv_peer));
                    vw.uleb(op.mv_lamport);
                    write_loro_value(vw, op.mark_val, keys);
                    break;
                case OP_UNKNOWN:
                    vk = op.raw_vkind;
                    if (op.mapval.k == Value::I64) vw.sleb(op.mapval.i);
                    else {
                        uint64_t bits; std::memcpy(&bits, &op.mapval.f, 8);
                        for (int i = 7; i >= 0; i--) vw.u8((uint8_t)(bits >> (8 * i)));
                    }
                    break;
            }
            c_cidx.push_back((int64_t)cidx);
            c_prop.push_back(prop);
            c_vt.push_back(vk);
            c_len.push_back((uint64_t)op.atom_len());
        }
    }
    // ContainerArena::from_containers (arena.rs:103-147): roots register keys, normals register peers
    Writer cw;
    cw.varint(cids.vec.size());
    for (auto& cid : cids.vec) {
        cw.varint(4);
        cw.u8(cid.root ? 1 : 0);
        cw.u8(cid.type);
        if (cid.root) {
            cw.varint(0);
            cw.zigzag((int64_t)keys.reg(cid.name));
        } else {
            cw.varint(peers.reg(cid.peer));
            cw.zigzag(cid.counter);
        }
    }
    Writer kw;
    for (auto& k : keys.vec) {
        kw.uleb(k.size());
        kw.bytes(k);
    }
    Writer pw;
    if (!positions.vec.empty()) {  // PositionArena::encode_v2 (arena.rs:218-224)
        Writer c0, c1;
        AnyRleEncoder<uint64_t, WrVar> e(c0, WrVar());
        c1.varint(positions.vec.size());
        std::string last;
        for (auto& p : positions.vec) {
            size_t common = 0;
            while (common < last.size() && common < p.size() && last[common] == p[common]) common++;
            e.append(common);
            c1.varint(p.size() - common);
            c1.bytes((const uint8_t*)p.data() + common, p.size() - common);
            last = p;
        }
        e.finish();
        columnar_write_wrapped(pw, {c0.buf, c1.buf});
    }
    Writer ow;
    {
        Writer a, b, c, d;
        delta_rle_encode(a, c_cidx);
        delta_rle_encode(b, c_prop);
        AnyRleEncoder<uint64_t, WrU8> e1(c, WrU8());
        for (auto x : c_vt) e1.append(x);
        e1.finish();
        AnyRleEncoder<uint64_t, WrVar> e2(d, WrVar());
        for (auto x : c_len) e2.append(x);
        e2.finish();
        columnar_write_wrapped(ow, {a.buf, b.buf, c.buf, d.buf});
    }
    Writer dw;
    if (!d_peer.empty()) {
        Writer a, b, c;
        delta_rle_encode(a, d_peer);
        delta_rle_encode(b, d_ctr);
        delta_rle_encode(c, d_len);
        columnar_write_wrapped(dw, {a.buf, b.buf, c.buf});
    }
    // encode_changes (block_meta_encode.rs:13-88)
    Writer lens, dep_self_w, dep_len_w, dep_peer_w, dep_ctr_w, lam_w, ts_w, msglen_w;
    std::string msgs;
    {
        std::vector<bool> dep_self;
        std::vector<int64_t> dep_ctrs, lams, tss;
        AnyRleEncoder<uint64_t, WrVar> dl(dep_len_w, WrVar()), dpi(dep_peer_w, WrVar()),
            ml(msglen_w, WrVar());
        for (size_t i = 0; i < block.size(); i++) {
            const Change& c = block[i];
            bool last = i + 1 == block.size();
            if (!last) {
                lens.uleb((uint64_t)c.atom_len());
                lams.push_back(c.lamport);
            }
            tss.push_back(c.timestamp);
            ml.append(c.has_msg ? c.msg.size() : 0);
            if (c.has_msg) msgs += c.msg;
            bool ds = false;
            size_t others = 0;
            for (auto& d : c.deps) {
                if (d.peer == peer)
                    ds = true;
                else {
                    dpi.append(peers.reg(d.peer));
                    dep_ctrs.push_back(d.counter);
                    others++;
                }
            }
            dep_self.push_back(ds);
            dl.append(others);
        }
        dl.finish();
        dpi.finish();
        ml.finish();
        bool_rle_encode(dep_self_w, dep_self);
        dod_encode(dep_ctr_w, dep_ctrs);
        dod_encode(lam_w, lams);
        dod_encode(ts_w, tss);
    }
    Writer hw;
    hw.uleb(peers.vec.size());
    for (auto p : peers.vec)
        for (int k = 0; k < 8; k++) hw.u8((uint8_t)(p >> (8 * k)));
    hw.bytes(lens.buf);
    hw.bytes(dep_self_w.buf);
    hw.bytes(dep_len_w.buf);
    hw.bytes(dep_peer_w.buf);
    hw.bytes(dep_ctr_w.buf);
    hw.bytes(lam_w.buf);
    Writer mw;
    mw.bytes(ts_w.buf);
    mw.bytes(msglen_w.buf);
    mw.bytes(msgs);

    Writer out;
    const Change& first = block.front();
    const Change& last = block.back();
    out.varint((uint32_t)first.id.counter);
    out.varint((uint32_t)(last.ctr_end() - first.id.counter));
    out.varint(first.lamport);
    out.varint(last.lamport_end() - first.lamport);
    out.varint(block.size());
    const std::vector<uint8_t>* secs[8] = {&hw.buf, &mw.buf, &cw.buf, &kw.buf,
                                          &pw.buf, &ow.buf, &dw.buf, &vw.buf};
    for (int i = 0; i < 8; i++) out.len_bytes(*secs[i]);
    if (out_secs)
        for (int i = 0; i < 8; i++) out_secs->sec[i] = *secs[i];
    return out.buf;
}

// ------------------------------------------------------------------ blob header / framing
enum : uint16_t { MODE_FAST_SNAPSHOT = 3, MODE_FAST_UPDATES = 4 };
enum BlobErr { BLOB_OK = 0, BLOB_TOO_SHORT, BLOB_BAD_MAGIC, BLOB_BAD_CHECKSUM, BLOB_BAD_MODE };

// parse_header_and_body + check_checksum (encoding.rs:278-330)
inline BlobErr parse_blob(const uint8_t* b, size_t n, uint16_t* mode, const uint8_t** body,
                          size_t* body_len) {
    if (n < 22) return BLOB_TOO_SHORT;
    if (std::memcmp(b, "loro", 4) != 0) return BLOB_BAD_MAGIC;
    uint16_t m = (uint16_t)((b[20] << 8) | b[21]);
    if (m != MODE_FAST_SNAPSHOT && m != MODE_FAST_UPDATES) return BLOB_BAD_MODE;
    uint32_t expect = rd32le(b + 16);
    if (xxh32(b + 20, n - 20, XXH_SEED_LORO) != expect) return BLOB_BAD_CHECKSUM;
    *mode = m;
    *body = b + 22;
    *body_len = n - 22;
    return BLOB_OK;
}
// encode_with (encoding.rs:397-416)
inline std::vector<uint8_t> wrap_blob(uint16_t mode, const std::vector<uint8_t>& body) {
    std::vector<uint8_t> out(22 + body.size(), 0);
    std::memcpy(out.data(), "loro", 4);
    out[20] = (uint8_t)(mode >> 8);
    out[21] = (uint8_t)mode;
    std::memcpy(out.data() + 22, body.data(), body.size());
    uint32_t h = xxh32(out.data() + 20, out.size() - 20, XXH_SEED_LORO);
    out[16] = (uint8_t)h; out[17] = (uint8_t)(h >> 8); out[18] = (uint8_t)(h >> 16); out[19] = (uint8_t)(h >> 24);
    return out;
}
// FastUpdates body -> block slices (fast_snapshot.rs:270-288)
inline std::vector<std::pair<const uint8_t*, size_t>> split_updates_body(const uint8_t* body,
                                                                         size_t n) {
    std::vector<std::pair<const uint8_t*, size_t>> out;
    Reader r(body, n);
    while (!r.empty()) {
        uint64_t len = r.uleb();
        out.push_back({r.take(len), (size_t)len});
    }
    return out;
}

}  // namespace lo
