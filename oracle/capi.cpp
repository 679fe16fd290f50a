// ORACLE (test infrastructure) -- C entry points for ctypes (tests/, smoke(), bench.py CPU legs).
// Never linked into the product library.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <sstream>
#include <thread>
#include "doc.hpp"

using namespace lo;

static char* dup_str(const std::string& s, size_t* len) {
    char* p = (char*)std::malloc(s.size() + 1);
    std::memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (len) *len = s.size();
    return p;
}
static uint8_t* dup_bytes(const std::vector<uint8_t>& v, size_t* len) {
    uint8_t* p = (uint8_t*)std::malloc(v.size() ? v.size() : 1);
    if (!v.empty()) std::memcpy(p, v.data(), v.size());
    *len = v.size();
    return p;
}
static Value mk_value(int kind, int64_t i, double f, const char* s, size_t slen) {
    Value v;
    switch (kind) {
        case 0: v.k = Value::Null; break;
        case 1: v.k = Value::True; break;
        case 2: v.k = Value::False; break;
        case 3: v = Value::i64(i); break;
        case 4: v = Value::f64(f); break;
        case 5: v = Value::str(std::string(s, slen)); break;
        case 6: v = Value::str(std::string(s, slen)); v.k = Value::Binary; break;
        case 9: v.k = Value::Container; v.cid.root = false; v.cid.type = (uint8_t)i; break;
        default: v.k = Value::Null;
    }
    return v;
}
static std::string ranges_json(const std::map<PeerID, std::pair<Counter, Counter>>& r) {
    std::string o = "{";
    bool first = true;
    for (auto& kv : r) {
        if (!first) o += ",";
        first = false;
        o += "\"" + std::to_string(kv.first) + "\":[" + std::to_string(kv.second.first) + "," +
             std::to_string(kv.second.second) + "]";
    }
    return o + "}";
}

extern "C" {

void lo_free(void* p) { std::free(p); }
void* lo_doc_new(uint64_t peer) { return new Doc(peer); }
void lo_doc_free(void* d) { delete (Doc*)d; }
void lo_set_peer(void* d, uint64_t peer) { ((Doc*)d)->commit(); ((Doc*)d)->peer = peer; }
int lo_get_container(void* d, const char* name, size_t nlen, int type) {
    return ((Doc*)d)->get_container(std::string(name, nlen), (uint8_t)type);
}
int lo_text_insert(void* d, int cidx, size_t pos, const char* s, size_t slen) {
    return ((Doc*)d)->text_insert(cidx, pos, std::string(s, slen)) ? 0 : -1;
}
int lo_list_insert(void* d, int cidx, size_t pos, size_t n, const int* kinds, const int64_t* ints,
                   const double* f64s, const char** strs, const size_t* slens) {
    std::vector<Value> vals;
    for (size_t i = 0; i < n; i++)
        vals.push_back(mk_value(kinds[i], ints ? ints[i] : 0, f64s ? f64s[i] : 0, strs ? strs[i] : "",
                                slens ? slens[i] : 0));
    return ((Doc*)d)->list_insert(cidx, pos, vals) ? 0 : -1;
}
// nested values from Python: tag byte 0 null, 1 true, 2 false, 3 i64 (8 bytes LE), 4 f64 (8 bytes LE), 5 str / 6 binary
// (u32 length + bytes), 7 list (u32 n + items), 8 map (u32 n + n x (u32 key length, key, value))
static Value parse_tagged(const uint8_t*& p) {
    Value v;
    uint8_t tag = *p++;
    auto u32_ = [&]() { uint32_t x; std::memcpy(&x, p, 4); p += 4; return x; };
    switch (tag) {
        case 0: v.k = Value::Null; break;
        case 1: v.k = Value::True; break;
        case 2: v.k = Value::False; break;
        case 3: { int64_t x; std::memcpy(&x, p, 8); p += 8; v = Value::i64(x); break; }
        case 4: { double x; std::memcpy(&x, p, 8); p += 8; v = Value::f64(x); break; }
        case 5: case 6: { uint32_t n = u32_(); v = Value::str(std::string((const char*)p, n)); if (tag == 6) v.k = Value::Binary; p += n; break; }
        case 7: { uint32_t n = u32_(); v.k = Value::List; for (uint32_t i = 0; i < n; i++) v.list.push_back(parse_tagged(p)); break; }
        case 8: {
            uint32_t n = u32_();
            v.k = Value::Map;
            for (uint32_t i = 0; i < n; i++) { uint32_t kl = u32_(); std::string key((const char*)p, kl); p += kl; v.map.push_back({key, parse_tagged(p)}); }
            break;
        }
        default: v.k = Value::Null;
    }
    return v;
}
int lo_list_insert_tagged(void* d, int cidx, size_t pos, size_t n, const uint8_t* buf) {
    std::vector<Value> vals;
    const uint8_t* p = buf;
    for (size_t i = 0; i < n; i++) vals.push_back(parse_tagged(p));
    return ((Doc*)d)->list_insert(cidx, pos, vals) ? 0 : -1;
}
int lo_map_set_tagged(void* d, int cidx, const char* key, size_t klen, const uint8_t* buf) {
    const uint8_t* p = buf;
    Value v = parse_tagged(p);
    return ((Doc*)d)->map_set(cidx, std::string(key, klen), &v) ? 0 : -1;
}
int lo_seq_delete(void* d, int cidx, size_t pos, size_t len) {
    return ((Doc*)d)->seq_delete(cidx, pos, len) ? 0 : -1;
}
int lo_seq_len(void* d, int cidx) {
    Doc* doc = (Doc*)d;
    doc->ensure_state();
    return (int)doc->cstate(cidx).ids.size();
}
int lo_map_set(void* d, int cidx, const char* key, size_t klen, int kind, int64_t i, double f,
               const char* s, size_t slen) {
    Value v = mk_value(kind, i, f, s, slen);
    return ((Doc*)d)->map_set(cidx, std::string(key, klen), &v) ? 0 : -1;
}
int lo_map_delete(void* d, int cidx, const char* key, size_t klen) {
    return ((Doc*)d)->map_set(cidx, std::string(key, klen), nullptr) ? 0 : -1;
}
// id of the child container created by the last list_insert / map_set with a Container value
int lo_child_container(void* d, uint64_t peer, int counter, int type) {
    ContainerID c;
    c.root = false;
    c.peer = peer;
    c.counter = counter;
    c.type = (uint8_t)type;
    return ((Doc*)d)->register_container(c);
}
int lo_next_counter(void* d) { return ((Doc*)d)->next_counter(); }
// local tree ops: a parent is (root flag, peer, counter); index < 0 appends.  0 ok, -1 rejected.
int lo_tree_create(void* d, int cidx, int parent_root, uint64_t ppeer, int pctr, int index, int* out_ctr) {
    ID out;
    if (!((Doc*)d)->tree_create(cidx, parent_root != 0, ID{ppeer, pctr}, index, &out)) return -1;
    if (out_ctr) *out_ctr = out.counter;
    return 0;
}
int lo_tree_move(void* d, int cidx, uint64_t tpeer, int tctr, int parent_root, uint64_t ppeer, int pctr, int index) {
    return ((Doc*)d)->tree_move(cidx, ID{tpeer, tctr}, parent_root != 0, ID{ppeer, pctr}, index) ? 0 : -1;
}
int lo_tree_delete(void* d, int cidx, uint64_t tpeer, int tctr) {
    return ((Doc*)d)->tree_delete(cidx, ID{tpeer, tctr}) ? 0 : -1;
}
int lo_tree_meta(void* d, uint64_t tpeer, int tctr) { return ((Doc*)d)->tree_meta(ID{tpeer, tctr}); }
// alive nodes of a tree container as "peer:ctr" pairs (workload generation): writes up to cap ids, returns the count
int lo_tree_nodes(void* d, int cidx, uint64_t* peers, int* ctrs, int cap) {
    Doc* doc = (Doc*)d;
    doc->ensure_state();
    int n = 0;
    for (auto& x : doc->cstate(cidx).tree) {
        if (x.deleted) continue;
        if (n < cap) { peers[n] = x.id.peer; ctrs[n] = x.id.counter; }
        n++;
    }
    return n;
}
void lo_commit(void* d) { ((Doc*)d)->commit(); }

int lo_export(void* d, size_t n_from, const uint64_t* peers, const int32_t* counters, uint8_t** out,
              size_t* len) {
    std::map<PeerID, Counter> from;
    for (size_t i = 0; i < n_from; i++) from[peers[i]] = counters[i];
    try {
        *out = dup_bytes(((Doc*)d)->export_updates(from), len);
    } catch (std::exception& e) {
        return -1;
    }
    return 0;
}
// returns 0 ok / blob error code; status_json = {"success":{peer:[a,b]},"pending":{...},"err":"..."}
int lo_import(void* d, const uint8_t* bytes, size_t n, char** status_json) {
    Doc::ImportStatus st;
    std::string err;
    int rc;
    try {
        rc = ((Doc*)d)->import(bytes, n, &st, &err);
    } catch (std::exception& e) {
        rc = 12;
        err = e.what();
    }
    if (status_json) {
        std::string j = "{\"success\":" + ranges_json(st.success) + ",\"pending\":" + ranges_json(st.pending) +
                        ",\"err\":";
        json_escape(j, err);
        j += "}";
        *status_json = dup_str(j, nullptr);
    }
    return rc;
}
char* lo_json(void* d, size_t* len) {
    try {
        return dup_str(((Doc*)d)->to_json(), len);
    } catch (std::exception& e) {
        return dup_str(std::string("!error: ") + e.what(), len);
    }
}
int lo_inconsistent_delete(void* d) { return ((Doc*)d)->inconsistent_delete ? 1 : 0; }
char* lo_vv_json(void* d) {
    Doc* doc = (Doc*)d;
    doc->commit();
    std::string o = "{";
    bool first = true;
    for (auto& kv : doc->vv) {
        if (!first) o += ",";
        first = false;
        o += "\"" + std::to_string(kv.first) + "\":" + std::to_string(kv.second);
    }
    o += "}";
    return dup_str(o, nullptr);
}
char* lo_frontiers_json(void* d) {
    Doc* doc = (Doc*)d;
    doc->commit();
    std::vector<ID> f = doc->frontiers;
    std::sort(f.begin(), f.end());
    std::string o = "[";
    for (size_t i = 0; i < f.size(); i++) {
        if (i) o += ",";
        o += "[\"" + std::to_string(f[i].peer) + "\"," + std::to_string(f[i].counter) + "]";
    }
    o += "]";
    return dup_str(o, nullptr);
}
int lo_pending_count(void* d) { return (int)((Doc*)d)->pending.size(); }
int64_t lo_len_ops(void* d) {
    Doc* doc = (Doc*)d;
    doc->commit();
    int64_t n = 0;
    for (auto& kv : doc->vv) n += kv.second;
    return n;
}

// ---------------------------------------------------------------- decode dump (golden-vector checks)
static void dump_value(std::string& o, const Value& v) {
    switch (v.k) {
        case Value::Null: o += "null"; break;
        case Value::True: o += "true"; break;
        case Value::False: o += "false"; break;
        case Value::I64: o += std::to_string(v.i); break;
        case Value::F64: json_f64(o, v.f); break;
        case Value::Str: json_escape(o, v.s); break;
        case Value::Binary: {
            o += "{\"binary\":[";
            for (size_t i = 0; i < v.s.size(); i++) { if (i) o += ","; o += std::to_string((unsigned char)v.s[i]); }
            o += "]}";
            break;
        }
        case Value::List:
            o += "[";
            for (size_t i = 0; i < v.list.size(); i++) { if (i) o += ","; dump_value(o, v.list[i]); }
            o += "]";
            break;
        case Value::Map: {
            o += "{";
            for (size_t i = 0; i < v.map.size(); i++) {
                if (i) o += ",";
                json_escape(o, v.map[i].first);
                o += ":";
                dump_value(o, v.map[i].second);
            }
            o += "}";
            break;
        }
        case Value::Container:
            o += "{\"container\":" + std::to_string(v.cid.type) + ",\"peer\":\"" + std::to_string(v.cid.peer) +
                 "\",\"counter\":" + std::to_string(v.cid.counter) + "}";
            break;
    }
}
static void dump_cid(std::string& o, const ContainerID& c) {
    o += "{\"root\":" + std::string(c.root ? "true" : "false") + ",\"type\":" + std::to_string(c.type);
    if (c.root) { o += ",\"name\":"; json_escape(o, c.name); }
    else o += ",\"peer\":\"" + std::to_string(c.peer) + "\",\"counter\":" + std::to_string(c.counter);
    o += "}";
}
static void dump_changes(std::string& o, const std::vector<Change>& chs, const Doc& doc) {
    o += "[";
    for (size_t i = 0; i < chs.size(); i++) {
        const Change& c = chs[i];
        if (i) o += ",";
        o += "{\"peer\":\"" + std::to_string(c.id.peer) + "\",\"counter\":" + std::to_string(c.id.counter) +
             ",\"lamport\":" + std::to_string(c.lamport) + ",\"timestamp\":" + std::to_string(c.timestamp) +
             ",\"deps\":[";
        for (size_t k = 0; k < c.deps.size(); k++) {
            if (k) o += ",";
            o += "[\"" + std::to_string(c.deps[k].peer) + "\"," + std::to_string(c.deps[k].counter) + "]";
        }
        o += "],\"msg\":";
        if (c.has_msg) json_escape(o, c.msg); else o += "null";
        o += ",\"ops\":[";
        for (size_t k = 0; k < c.ops.size(); k++) {
            const Op& op = c.ops[k];
            if (k) o += ",";
            o += "{\"counter\":" + std::to_string(op.counter) + ",\"container\":";
            dump_cid(o, doc.containers[(size_t)op.cidx]);
            o += ",\"prop\":" + std::to_string(op.prop) + ",\"len\":" + std::to_string(op.atom_len()) + ",\"kind\":";
            switch (op.kind) {
                case OP_LIST_INSERT: {
                    o += "\"insert\",\"values\":[";
                    for (size_t j = 0; j < op.values.size(); j++) { if (j) o += ","; dump_value(o, op.values[j]); }
                    o += "]";
                    break;
                }
                case OP_TEXT_INSERT: o += "\"insert_text\",\"text\":"; json_escape(o, op.text); break;
                case OP_DELETE:
                    o += "\"delete\",\"id_start\":[\"" + std::to_string(op.del_start.peer) + "\"," +
                         std::to_string(op.del_start.counter) + "],\"signed_len\":" + std::to_string(op.del_len);
                    break;
                case OP_MAP_SET: o += "\"map_set\",\"key\":"; json_escape(o, op.key); o += ",\"value\":"; dump_value(o, op.mapval); break;
                case OP_MAP_DEL: o += "\"map_del\",\"key\":"; json_escape(o, op.key); break;
                case OP_TREE_CREATE: case OP_TREE_MOVE: case OP_TREE_DELETE:
                    o += op.kind == OP_TREE_CREATE ? "\"tree_create\"" : op.kind == OP_TREE_MOVE ? "\"tree_move\"" : "\"tree_delete\"";
                    o += ",\"target\":[\"" + std::to_string(op.target.peer) + "\"," + std::to_string(op.target.counter) + "]";
                    if (op.kind != OP_TREE_DELETE) {
                        o += ",\"parent\":";
                        if (op.parent_null) o += "null";
                        else o += "[\"" + std::to_string(op.parent.peer) + "\"," + std::to_string(op.parent.counter) + "]";
                        o += ",\"position\":[";
                        for (size_t j = 0; j < op.position.size(); j++) { if (j) o += ","; o += std::to_string((unsigned char)op.position[j]); }
                        o += "]";
                    }
                    break;
                case OP_STYLE_START: o += "\"style_start\",\"mark_len\":" + std::to_string(op.mark_len) + ",\"key\":"; json_escape(o, op.mark_key); break;
                case OP_STYLE_END: o += "\"style_end\""; break;
                case OP_LIST_MOVE: o += "\"list_move\""; break;
                case OP_LIST_SET: o += "\"list_set\""; break;
                default: o += "\"unknown\"";
            }
            o += "}";
        }
        o += "]}";
    }
    o += "]";
}
// Decode a FastUpdates blob (or, with raw_block!=0, one bare change block) and describe it as JSON.
char* lo_decode_dump(const uint8_t* bytes, size_t n, int raw_block, size_t* len) {
    Doc doc(0);
    std::string o;
    try {
        std::vector<std::pair<const uint8_t*, size_t>> blocks;
        uint16_t mode = 0;
        if (raw_block) {
            blocks.push_back({bytes, n});
        } else {
            const uint8_t* body;
            size_t blen;
            BlobErr e = parse_blob(bytes, n, &mode, &body, &blen);
            if (e != BLOB_OK) return dup_str("{\"error\":\"blob header " + std::to_string((int)e) + "\"}", len);
            if (mode != MODE_FAST_UPDATES) return dup_str("{\"error\":\"mode " + std::to_string(mode) + "\"}", len);
            blocks = split_updates_body(body, blen);
        }
        o = "{\"mode\":" + std::to_string(mode) + ",\"blocks\":[";
        for (size_t b = 0; b < blocks.size(); b++) {
            BlockMeta m;
            std::vector<Change> chs = decode_block(blocks[b].first, blocks[b].second, doc, &m);
            if (b) o += ",";
            o += "{\"counter_start\":" + std::to_string(m.counter_start) + ",\"counter_len\":" + std::to_string(m.counter_len) +
                 ",\"lamport_start\":" + std::to_string(m.lamport_start) + ",\"lamport_len\":" + std::to_string(m.lamport_len) +
                 ",\"n_changes\":" + std::to_string(m.n_changes) + ",\"peers\":[";
            for (size_t i = 0; i < m.peers.size(); i++) { if (i) o += ","; o += "\"" + std::to_string(m.peers[i]) + "\""; }
            o += "],\"keys\":[";
            for (size_t i = 0; i < m.keys.size(); i++) { if (i) o += ","; json_escape(o, m.keys[i]); }
            o += "],\"section_lens\":[";
            for (int i = 0; i < 8; i++) { if (i) o += ","; o += std::to_string(m.sec_len[i]); }
            o += "],\"changes\":";
            dump_changes(o, chs, doc);
            o += "}";
        }
        o += "]}";
    } catch (std::exception& e) {
        o = "{\"error\":";
        json_escape(o, e.what());
        o += "}";
    }
    return dup_str(o, len);
}
// decode one bare change block and re-encode it (encoder pin against golden blocks);
// section != -1 returns just that section (0..7) of the re-encoding
int lo_block_roundtrip(const uint8_t* bytes, size_t n, int section, uint8_t** out, size_t* len) {
    Doc doc(0);
    try {
        std::vector<Change> chs = decode_block(bytes, n, doc);
        EncodedSections secs;
        std::vector<uint8_t> re = encode_block(chs, doc, &secs);
        if (section >= 0 && section < 8) *out = dup_bytes(secs.sec[section], len);
        else *out = dup_bytes(re, len);
    } catch (std::exception& e) {
        std::string s = e.what();
        std::vector<uint8_t> v(s.begin(), s.end());
        *out = dup_bytes(v, len);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- codec primitive hooks
// values travel as little-endian i64 arrays
static std::vector<int64_t> rd_i64s(const uint8_t* in, size_t n) {
    std::vector<int64_t> v(n / 8);
    std::memcpy(v.data(), in, v.size() * 8);
    return v;
}
static std::vector<uint8_t> wr_i64s(const std::vector<int64_t>& v) {
    std::vector<uint8_t> o(v.size() * 8);
    std::memcpy(o.data(), v.data(), o.size());
    return o;
}
int lo_codec(const char* op, const uint8_t* in, size_t n, int64_t arg, uint8_t** out, size_t* len) {
    std::string name(op);
    try {
        Reader r(in, n);
        Writer w;
        std::vector<int64_t> vals;
        if (name == "str_arena_gens") {   // generation of the string arena buffer after each allocation (doc.hpp alloc_str)
            Doc d(1);
            std::vector<int64_t> lens = rd_i64s(in, n);
            for (int64_t l : lens) {
                Op op;
                op.text.assign((size_t)l, 'x');
                op.unicode_len = (uint32_t)l;
                d.alloc_str(op);
                vals.push_back(op.arena_gen);
            }
            *out = dup_bytes(wr_i64s(vals), len);
        } else if (name == "f64_json") {   // serde_json text of a double (doc.hpp json_f64): checker for the device formatter
            double d;
            std::memcpy(&d, in, 8);
            std::string o;
            json_f64(o, d);
            *out = dup_bytes(std::vector<uint8_t>(o.begin(), o.end()), len);
        } else if (name == "xxh32") {
            vals.push_back(xxh32(in, n, (uint32_t)arg));
            *out = dup_bytes(wr_i64s(vals), len);
        } else if (name == "varint_dec") { vals.push_back((int64_t)r.varint()); vals.push_back((int64_t)(n - r.remaining())); *out = dup_bytes(wr_i64s(vals), len); }
        else if (name == "zigzag_dec") { vals.push_back(r.zigzag()); vals.push_back((int64_t)(n - r.remaining())); *out = dup_bytes(wr_i64s(vals), len); }
        else if (name == "sleb_dec") { vals.push_back(r.sleb()); vals.push_back((int64_t)(n - r.remaining())); *out = dup_bytes(wr_i64s(vals), len); }
        else if (name == "varint_enc") { w.varint((uint64_t)arg); *out = dup_bytes(w.buf, len); }
        else if (name == "zigzag_enc") { w.zigzag(arg); *out = dup_bytes(w.buf, len); }
        else if (name == "sleb_enc") { w.sleb(arg); *out = dup_bytes(w.buf, len); }
        else if (name == "boolrle_dec") {
            std::vector<bool> b;
            bool state = false;
            while (!r.empty()) { uint64_t l = r.varint(); for (uint64_t i = 0; i < l; i++) b.push_back(state); state = !state; }
            for (bool x : b) vals.push_back(x);
            *out = dup_bytes(wr_i64s(vals), len);
        } else if (name == "boolrle_enc") {
            std::vector<bool> b;
            for (auto x : rd_i64s(in, n)) b.push_back(x != 0);
            bool_rle_encode(w, b);
            *out = dup_bytes(w.buf, len);
        } else if (name == "anyrle_u8_dec" || name == "anyrle_u32_dec") {
            std::vector<uint64_t> v;
            if (name == "anyrle_u8_dec") any_rle_decode_all<uint64_t>(r, v, RdU8());
            else any_rle_decode_all<uint64_t>(r, v, RdVar());
            for (auto x : v) vals.push_back((int64_t)x);
            *out = dup_bytes(wr_i64s(vals), len);
        } else if (name == "anyrle_u8_enc" || name == "anyrle_u32_enc") {
            if (name == "anyrle_u8_enc") { AnyRleEncoder<uint64_t, WrU8> e(w, WrU8()); for (auto x : rd_i64s(in, n)) e.append((uint64_t)x); e.finish(); }
            else { AnyRleEncoder<uint64_t, WrVar> e(w, WrVar()); for (auto x : rd_i64s(in, n)) e.append((uint64_t)x); e.finish(); }
            *out = dup_bytes(w.buf, len);
        } else if (name == "anyrle_i32_dec") {
            std::vector<int64_t> v;
            any_rle_decode_all<int64_t>(r, v, RdZig());
            *out = dup_bytes(wr_i64s(v), len);
        } else if (name == "deltarle_dec") { *out = dup_bytes(wr_i64s(delta_rle_decode_all(r)), len); }
        else if (name == "deltarle_enc") { delta_rle_encode(w, rd_i64s(in, n)); *out = dup_bytes(w.buf, len); }
        else if (name == "dod_dec") {  // arg = n values; appends consumed byte count as last value
            std::vector<int64_t> v = dod_take_n(r, (size_t)arg);
            v.push_back((int64_t)(n - r.remaining()));
            *out = dup_bytes(wr_i64s(v), len);
        } else if (name == "dod_enc") { dod_encode(w, rd_i64s(in, n)); *out = dup_bytes(w.buf, len); }
        else return -2;
    } catch (std::exception& e) {
        std::string s = e.what();
        std::vector<uint8_t> v(s.begin(), s.end());
        *out = dup_bytes(v, len);
        return -1;
    }
    return 0;
}

// ---------------------------------------------------------------- batch import (CPU baseline leg)
// Imports docs[i] = blobs[offsets[i]..offsets[i+1]) each into a fresh doc on `threads` host threads;
// flags bit0: also compute deep JSON, bit1: also re-export all updates.  Returns total atom ops merged;
// *out_hash = xor of xxh32 over each doc's JSON (order independent), *seconds = wall time.
int64_t lo_bench_import(const uint8_t* blobs, const uint64_t* offsets, size_t n_docs, int threads, int flags,
                        uint64_t* out_hash, double* seconds) {
    std::atomic<size_t> next(0);
    std::atomic<int64_t> ops(0);
    std::atomic<uint64_t> hash(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        while (true) {
            size_t i = next.fetch_add(1);
            if (i >= n_docs) break;
            Doc doc(0);
            Doc::ImportStatus st;
            int rc = doc.import(blobs + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), &st);
            if (rc != 0) continue;
            int64_t n = 0;
            for (auto& kv : doc.vv) n += kv.second;
            ops += n;
            if (flags & 1) {
                std::string j = doc.to_json();
                hash ^= ((uint64_t)xxh32((const uint8_t*)j.data(), j.size(), 0) << 32) | (uint64_t)j.size();
            } else
                doc.ensure_state();
            if (flags & 2) {
                std::vector<uint8_t> e = doc.export_updates({});
                hash ^= xxh32(e.data(), e.size(), 1);
            }
        }
    };
    std::vector<std::thread> ts;
    for (int t = 1; t < threads; t++) ts.emplace_back(work);
    work();
    for (auto& t : ts) t.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (out_hash) *out_hash = hash.load();
    return ops.load();
}

}  // extern "C"
