// Host-side zone builder: snapshot JSON-lines -> bb::ZoneImage.
//
// Replaces the ingest half of lib/zk.js (TreeNode ctor :78-97, onChildrenChanged :120-138,
// onDataChanged :139-194) for a whole-subtree snapshot, and pre-evaluates the query-independent
// parts of lib/server.js resolve()/resolvePtr() per node (see zone_image.h).
//
// JSON handling is a single-pass "tape" parser (no per-value allocation): 10M-znode
// snapshots (~1.1 GB of text) build in seconds.
#include "zone_image.h"
#include "../../include/binder_b200.h"

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

using namespace bb;

// ---------------------------------------------------------------------------------------
// JSON tape
// ---------------------------------------------------------------------------------------
enum JT : uint8_t { J_NULL, J_FALSE, J_TRUE, J_NUM, J_STR, J_ARR, J_OBJ };
struct Tok {
    uint8_t  type;
    uint32_t a, b;       // J_STR: [a, a+b) in pool.  J_ARR/J_OBJ: b = number of children (pairs for OBJ)
    uint32_t next;       // index of the token after this value's subtree
    double   num;
};
struct Tape {
    std::vector<Tok> t;
    std::string pool;
    const char* p; const char* e;

    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    static void put_utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o.push_back((char)c);
        else if (c < 0x800) { o.push_back((char)(0xC0 | c >> 6)); o.push_back((char)(0x80 | (c & 63))); }
        else if (c < 0x10000) { o.push_back((char)(0xE0 | c >> 12)); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
        else { o.push_back((char)(0xF0 | c >> 18)); o.push_back((char)(0x80 | ((c >> 12) & 63))); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4) return false;
        v = 0;
        for (int i = 0; i < 4; i++) {
            unsigned c = (unsigned char)p[i], d;
            if (c - '0' < 10) d = c - '0'; else if ((c | 32) - 'a' < 6) d = (c | 32) - 'a' + 10; else return false;
            v = v << 4 | d;
        }
        p += 4; return true;
    }
    bool str(Tok& k) {
        ++p;                                    // opening quote
        k.type = J_STR; k.a = (uint32_t)pool.size();
        for (;;) {
            const char* s = p;
            while (p < e && *p != '"' && *p != '\\' && (unsigned char)*p >= 0x20) ++p;
            pool.append(s, p - s);
            if (p >= e || (unsigned char)*p < 0x20) return false;
            if (*p == '"') { ++p; break; }
            ++p; if (p >= e) return false;
            char x = *p++;
            switch (x) {
            case '"': case '\\': case '/': pool.push_back(x); break;
            case 'b': pool.push_back('\b'); break; case 'f': pool.push_back('\f'); break;
            case 'n': pool.push_back('\n'); break; case 'r': pool.push_back('\r'); break;
            case 't': pool.push_back('\t'); break;
            case 'u': {
                uint32_t v; if (!hex4(v)) return false;
                if (v >= 0xD800 && v < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                    const char* save = p; p += 2; uint32_t lo;
                    if (hex4(lo) && lo >= 0xDC00 && lo < 0xE000) v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00);
                    else p = save;
                }
                put_utf8(pool, v); break; }
            default: return false;
            }
        }
        k.b = (uint32_t)pool.size() - k.a;
        return true;
    }
    bool num(Tok& k) {
        const char* s = p;
        if (p < e && *p == '-') ++p;
        if (p >= e) return false;
        if (*p == '0') ++p;
        else if (*p >= '1' && *p <= '9') { while (p < e && (unsigned)(*p - '0') < 10) ++p; }
        else return false;
        if (p < e && *p == '.') { ++p; if (p >= e || (unsigned)(*p - '0') >= 10) return false; while (p < e && (unsigned)(*p - '0') < 10) ++p; }
        if (p < e && (*p == 'e' || *p == 'E')) {
            ++p; if (p < e && (*p == '+' || *p == '-')) ++p;
            if (p >= e || (unsigned)(*p - '0') >= 10) return false;
            while (p < e && (unsigned)(*p - '0') < 10) ++p;
        }
        char buf[64]; size_t n = (size_t)(p - s);
        k.type = J_NUM;
        if (n < sizeof buf) { memcpy(buf, s, n); buf[n] = 0; k.num = strtod(buf, nullptr); }
        else { std::string tmp(s, n); k.num = strtod(tmp.c_str(), nullptr); }
        return true;
    }
    bool lit(const char* w, size_t n) { if ((size_t)(e - p) < n || memcmp(p, w, n)) return false; p += n; return true; }
    bool value(int depth) {
        if (depth > 200) return false;
        ws(); if (p >= e) return false;
        size_t me = t.size(); t.emplace_back();
        Tok k; k.a = k.b = k.next = 0; k.num = 0;
        bool ok = true;
        switch (*p) {
        case '{': {
            ++p; k.type = J_OBJ; ws();
            if (p < e && *p == '}') { ++p; break; }
            for (;;) {
                ws(); if (p >= e || *p != '"') return false;
                size_t ki = t.size(); t.emplace_back(); Tok kk; kk.next = 0; kk.num = 0;
                if (!str(kk)) return false;
                kk.next = (uint32_t)t.size(); t[ki] = kk;
                ws(); if (p >= e || *p != ':') return false; ++p;
                if (!value(depth + 1)) return false;
                ++k.b; ws(); if (p >= e) return false;
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; break; }
                return false;
            }
            break; }
        case '[': {
            ++p; k.type = J_ARR; ws();
            if (p < e && *p == ']') { ++p; break; }
            for (;;) {
                if (!value(depth + 1)) return false;
                ++k.b; ws(); if (p >= e) return false;
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; break; }
                return false;
            }
            break; }
        case '"': ok = str(k); break;
        case 't': k.type = J_TRUE; ok = lit("true", 4); break;
        case 'f': k.type = J_FALSE; ok = lit("false", 5); break;
        case 'n': k.type = J_NULL; ok = lit("null", 4); break;
        default: ok = num(k);
        }
        if (!ok) return false;
        k.next = (uint32_t)t.size();
        t[me] = k;
        return true;
    }
    // parse [b, e) as one JSON document appended to the tape; returns root index or -1
    int parse(const char* b, const char* end) {
        size_t t0 = t.size(), p0 = pool.size();
        p = b; e = end;
        bool ok = value(0);
        if (ok) { ws(); ok = (p == e); }
        if (!ok) { t.resize(t0); pool.resize(p0); return -1; }
        return (int)t0;
    }
    void clear() { t.clear(); pool.clear(); }
    bool is_obj(int v) const { return v >= 0 && (t[v].type == J_OBJ || t[v].type == J_ARR); }   // typeof 'object', non-null
    // own-property read; JSON.parse keeps the last duplicate
    int get(int v, const char* key, size_t klen) const {
        if (v < 0 || t[v].type != J_OBJ) return -1;
        int found = -1; uint32_t i = (uint32_t)v + 1;
        for (uint32_t c = 0; c < t[v].b; c++) {
            const Tok& k = t[i];
            uint32_t val = i + 1;
            if (k.b == klen && memcmp(pool.data() + k.a, key, klen) == 0) found = (int)val;
            i = t[val].next;
        }
        return found;
    }
    int get(int v, const char* key) const { return get(v, key, strlen(key)); }
    bool str_eq(int v, const char* s) const { return v >= 0 && t[v].type == J_STR && t[v].b == strlen(s) && memcmp(pool.data() + t[v].a, s, t[v].b) == 0; }
};

// ---------------------------------------------------------------------------------------
// contract predicates (DESIGN.md "Contract": what mname's record constructors accept)
// ---------------------------------------------------------------------------------------
bool uint_ok(const Tape& J, int v, double limit, uint32_t& out) {
    if (v < 0 || J.t[v].type != J_NUM) return false;
    double d = J.t[v].num;
    if (!(d >= 0.0) || !(d < limit) || std::floor(d) != d) return false;
    out = (uint32_t)d; return true;
}
bool ipv4_ok(const char* s, size_t n, uint32_t& out) {
    size_t i = 0; uint32_t acc = 0;
    for (int o = 0; o < 4; o++) {
        size_t st = i; unsigned v = 0;
        while (i < n && (unsigned)(s[i] - '0') < 10 && i - st < 4) { v = v * 10 + (unsigned)(s[i] - '0'); ++i; }
        size_t nd = i - st;
        if (nd < 1 || nd > 3 || v > 255 || (nd > 1 && s[st] == '0')) return false;
        acc = acc << 8 | v;
        if (o < 3) { if (i >= n || s[i] != '.') return false; ++i; }
    }
    if (i != n) return false;
    out = acc; return true;
}
// hostname of scheme://[userinfo@]host[:port][/...]  (url.parse at lib/server.js:297-298)
bool url_host_ipv4(const char* s, size_t n, uint32_t& out) {
    size_t i = 0;
    if (n == 0 || !(((unsigned)(s[0] | 32) - 'a') < 26)) return false;
    while (i < n && ((((unsigned)(s[i] | 32) - 'a') < 26) || (unsigned)(s[i] - '0') < 10 || s[i] == '+' || s[i] == '.' || s[i] == '-')) ++i;
    if (n - i < 3 || memcmp(s + i, "://", 3)) return false;
    i += 3;
    size_t e = i;
    while (e < n && s[e] != '/' && s[e] != '?' && s[e] != '#') ++e;
    for (size_t k = e; k > i; k--) if (s[k - 1] == '@') { i = k; break; }
    for (size_t k = e; k > i; k--) if (s[k - 1] == ':') { e = k - 1; break; }
    return ipv4_ok(s + i, e - i, out);          // digits and dots: lower-casing is a no-op
}

const char* const kHostLike[] = { "db_host", "host", "load_balancer", "moray_host", "redis_host", "ops_host", "rr_host" };
const char* const kSvcKid[] = { "load_balancer", "moray_host", "ops_host", "rr_host", "redis_host" };

// ---------------------------------------------------------------------------------------
// flattened nodes
// ---------------------------------------------------------------------------------------
enum : uint16_t {
    NF_KIDTYPE = 1,      // data.type passes the service child filter (lib/server.js:352-360)
    NF_SUB_OBJ = 2,      // data[data.type] is a non-null object
    NF_HAS_TTL = 4,      // data.ttl or data[type].ttl defined
    NF_TTL_OK = 8,
    NF_ADDR_NULL = 16,
    NF_ADDR_OK = 32,
    NF_PORTS_LIST = 64,  // ports is a non-empty list of valid ports (stored in `ports`)
    NF_PORTS_BAD = 128,  // ports present but unusable
    NF_REV = 256,        // registers in ca_revLookup
    NF_HOSTLIKE = 512,
};
struct Node {
    uint32_t parent, name_off, name_len;
    uint32_t first_kid = 0, last_kid = 0, next_sib = 0;      // 0 = none (node 0 is the root)
    uint32_t ttl = 30, addr = 0, extra = 0xFFFFFFFFu;        // extra: index into svcs / ports
    uint32_t rev_off = 0, rev_len = 0;                        // address string in pool (reverse key)
    uint16_t flags = 0; uint8_t kind = K_INVALID; uint8_t pad = 0;
};
struct SvcInfo { std::string srvce, proto; bool has_srvce = false, has_proto = false, port_ok = false; uint32_t port = 0; uint32_t ttl = 30; };

struct Builder {
    std::string dns_domain;                 // lower-cased root domain
    std::vector<Node> nodes;
    std::string pool;                       // names + address strings
    std::vector<SvcInfo> svcs;
    std::vector<std::vector<uint16_t>> ports;
    std::vector<uint32_t> child_tab;        // open-addressed (parent,name) -> node index + 1
    uint32_t child_mask = 0;
    Tape J;

    static uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
    uint64_t child_hash(uint32_t parent, const char* s, size_t n) const {
        uint64_t h = 0x9E3779B97F4A7C15ULL ^ parent;
        for (size_t i = 0; i < n; i++) h = (h ^ (unsigned char)s[i]) * 0x100000001B3ULL;
        return mix(h);
    }
    void child_grow() {
        uint32_t cap = child_tab.empty() ? 1024 : (uint32_t)child_tab.size() * 2;
        std::vector<uint32_t> nt(cap, 0);
        uint32_t m = cap - 1;
        for (uint32_t v : child_tab) if (v) {
            const Node& nd = nodes[v - 1];
            uint64_t h = child_hash(nd.parent, pool.data() + nd.name_off, nd.name_len);
            uint32_t i = (uint32_t)h & m; while (nt[i]) i = (i + 1) & m; nt[i] = v;
        }
        child_tab.swap(nt); child_mask = m;
    }
    int find_child(uint32_t parent, const char* s, size_t n) const {
        if (child_tab.empty()) return -1;
        uint32_t i = (uint32_t)child_hash(parent, s, n) & child_mask;
        while (child_tab[i]) {
            const Node& nd = nodes[child_tab[i] - 1];
            if (nd.parent == parent && nd.name_len == n && memcmp(pool.data() + nd.name_off, s, n) == 0) return (int)(child_tab[i] - 1);
            i = (i + 1) & child_mask;
        }
        return -1;
    }
    uint32_t add_node(uint32_t parent, const char* s, size_t n) {
        Node nd; nd.parent = parent; nd.name_off = (uint32_t)pool.size(); nd.name_len = (uint32_t)n;
        pool.append(s, n);
        uint32_t id = (uint32_t)nodes.size();
        nodes.push_back(nd);
        if (id != 0) {
            if ((nodes.size() + 1) * 2 > child_tab.size()) child_grow();
            uint32_t i = (uint32_t)child_hash(parent, s, n) & child_mask;
            while (child_tab[i]) i = (i + 1) & child_mask;
            child_tab[i] = id + 1;
            Node& p = nodes[parent];
            if (p.last_kid) nodes[p.last_kid].next_sib = id; else p.first_kid = id;
            p.last_kid = id;
        }
        return id;
    }

    // lib/zk.js:139-194 + the query-independent half of lib/server.js:249-274,296-332
    void ingest(uint32_t id, int v) {
        Node& nd = nodes[id];
        nd.kind = K_INVALID; nd.flags = 0; nd.ttl = 30;
        if (v < 0 || J.t[v].type != J_OBJ) return;           // null, array (no .type), or nothing stored
        int type = J.get(v, "type");
        if (type < 0 || J.t[type].type != J_STR) return;
        const char* ts = J.pool.data() + J.t[type].a; size_t tl = J.t[type].b;
        int sub = J.get(v, ts, tl);
        bool hostlike = false, kidtype = false;
        for (auto h : kHostLike) if (strlen(h) == tl && !memcmp(h, ts, tl)) hostlike = true;
        for (auto h : kSvcKid) if (strlen(h) == tl && !memcmp(h, ts, tl)) kidtype = true;
        if (kidtype) nd.flags |= NF_KIDTYPE;
        if (hostlike) nd.flags |= NF_HOSTLIKE;
        if (!J.is_obj(sub)) return;                           // :251-253 / :366-368
        nd.flags |= NF_SUB_OBJ;
        // ttl: record.ttl then record[type].ttl (:270-274); kept even when unusable so that a
        // service child can tell "has its own ttl" from "inherits" (:389-393)
        int tv = -1, a = J.get(v, "ttl"), b = J.get(sub, "ttl");
        if (a >= 0) tv = a;
        if (b >= 0) tv = b;
        bool ttl_ok = true;
        if (tv >= 0) { nd.flags |= NF_HAS_TTL; ttl_ok = uint_ok(J, tv, 2147483648.0, nd.ttl); }
        if (ttl_ok) nd.flags |= NF_TTL_OK;
        bool is_service = tl == 7 && !memcmp(ts, "service", 7);
        bool is_db = tl == 8 && !memcmp(ts, "database", 8);
        if (hostlike) {
            int ad = J.get(sub, "address");
            if (ad >= 0 && J.t[ad].type == J_NULL) nd.flags |= NF_ADDR_NULL;
            if (ad >= 0 && J.t[ad].type == J_STR) {
                const char* as = J.pool.data() + J.t[ad].a; size_t al = J.t[ad].b;
                if (ipv4_ok(as, al, nd.addr)) nd.flags |= NF_ADDR_OK;
                if (al > 0) {                                  // lib/zk.js:183-188 (strings only)
                    nd.flags |= NF_REV; nd.rev_off = (uint32_t)pool.size(); nd.rev_len = (uint32_t)al;
                    pool.append(as, al);
                }
            }
            int pv = J.get(sub, "ports");
            if (pv >= 0 && !(J.t[pv].type == J_ARR && J.t[pv].b == 0)) {     // :383-385
                bool ok = J.t[pv].type == J_ARR;
                std::vector<uint16_t> pl;
                if (ok) {
                    uint32_t i = (uint32_t)pv + 1;
                    for (uint32_t c = 0; c < J.t[pv].b; c++) { uint32_t p; if (uint_ok(J, (int)i, 65536.0, p)) pl.push_back((uint16_t)p); else ok = false; i = J.t[i].next; }
                    if (pl.size() > 255) ok = false;
                }
                if (ok) { nd.flags |= NF_PORTS_LIST; nd.extra = (uint32_t)ports.size(); ports.push_back(std::move(pl)); }
                else nd.flags |= NF_PORTS_BAD;
            }
        }
        if (!ttl_ok) return;                                  // K_INVALID (contract)
        if (hostlike) { nd.kind = (nd.flags & NF_ADDR_OK) ? K_ADDR : K_ADDR_BAD; return; }
        if (is_db) {
            int pr = J.get(sub, "primary");
            nd.kind = (pr >= 0 && J.t[pr].type == J_STR && url_host_ipv4(J.pool.data() + J.t[pr].a, J.t[pr].b, nd.addr)) ? K_ADDR : K_ADDR_BAD;
            return;
        }
        if (is_service) {
            int s = sub;
            int inner = J.get(s, "service");
            if (inner >= 0 && J.t[inner].type == J_NULL) return;             // contract: null.ttl throws
            if (J.is_obj(inner)) s = inner;                                   // :324-325
            SvcInfo si; si.ttl = nd.ttl;
            int st = J.get(s, "ttl");
            if (st >= 0 && !uint_ok(J, st, 2147483648.0, si.ttl)) return;     // :331-332 (contract)
            int sv = J.get(s, "srvce"), pr = J.get(s, "proto"), po = J.get(s, "port");
            if (sv >= 0 && J.t[sv].type == J_STR) { si.has_srvce = true; si.srvce.assign(J.pool.data() + J.t[sv].a, J.t[sv].b); }
            if (pr >= 0 && J.t[pr].type == J_STR) { si.has_proto = true; si.proto.assign(J.pool.data() + J.t[pr].a, J.t[pr].b); }
            si.port_ok = uint_ok(J, po, 65536.0, si.port);
            nd.kind = K_SERVICE; nd.extra = (uint32_t)svcs.size(); svcs.push_back(std::move(si));
            return;
        }
        nd.kind = K_UNKNOWN;
    }

    // lower-cased fqdn of a node (lib/zk.js:80-84): name + '.' + parent's domain
    void domain_of(uint32_t id, std::string& out) const {
        out.clear();
        for (uint32_t cur = id;; cur = nodes[cur].parent) {
            if (cur == 0) { out += dns_domain; break; }
            size_t at = out.size();
            out.append(pool.data() + nodes[cur].name_off, nodes[cur].name_len);
            for (size_t i = at; i < out.size(); i++) if (out[i] >= 'A' && out[i] <= 'Z') out[i] = (char)(out[i] + 32);
            out.push_back('.');
        }
    }
};

// dotted name -> wire labels (no terminator); false if a label is empty or > 63
bool to_wire(const char* s, size_t n, std::string& out) {
    out.clear();
    size_t st = 0;
    for (size_t i = 0; i <= n; i++) {
        if (i == n || s[i] == '.') {
            size_t l = i - st;
            if (l < 1 || l > 63) return false;
            out.push_back((char)l); out.append(s + st, l);
            st = i + 1;
        }
    }
    return true;
}

struct TableBuilder {
    ZoneImage* z;
    std::vector<uint8_t> arena;
    uint32_t mask;
    uint32_t arena_put(const void* p, size_t n) {
        while (arena.size() & 3) arena.push_back(0);
        uint32_t off = (uint32_t)arena.size();
        arena.insert(arena.end(), (const uint8_t*)p, (const uint8_t*)p + n);
        return off;
    }
    bool key_eq(const Slot& s, uint32_t ns, const uint8_t* k, uint32_t len) const {
        if ((uint32_t)(s.ns & 1) != ns) return false;
        if (len <= KEY_INLINE_MAX) return s.klen == len && memcmp(s.key, k, len) == 0;
        if (s.klen != KLEN_OVERFLOW) return false;
        uint32_t off, l; memcpy(&off, s.key, 4); memcpy(&l, s.key + 4, 4);
        return l == len && memcmp(arena.data() + off, k, len) == 0;
    }
    uint32_t nranks = 1, rank = 0;
    bool mine(uint32_t ns, const uint8_t* k, uint32_t len) const { return nranks == 1 || owner_of(hash_key(ns, k, len), nranks) == rank; }
    bool failed = false;         // a cuckoo insertion ran out of kicks: the caller rebuilds with a larger table
    std::vector<uint32_t> h2s;   // second hash of the key resident in each slot (host-side only: evictions need it)
    void fill(Slot& s, uint32_t h, uint32_t ns, const uint8_t* k, uint32_t len, uint8_t kind, uint32_t ttl, uint32_t val) {
        memset(&s, 0, sizeof s);
        s.hash = h;
        uint32_t dots = 0; bool clean = true;
        for (uint32_t i = 0; i < len; i++) {
            const uint8_t c = k[i];
            dots += c == '.';
            if (!((c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == '.')) clean = false;
        }
        s.ns = (uint8_t)(ns | (ns == NS_FORWARD ? (dots & 127) << 1 : 0));
        s.flags = (ns == NS_FORWARD && clean) ? SLOT_KEY_CLEAN : 0;
        if (len <= KEY_INLINE_MAX) { s.klen = (uint8_t)len; memcpy(s.key, k, len); }
        else { s.klen = KLEN_OVERFLOW; uint32_t off = arena_put(k, len); memcpy(s.key, &off, 4); memcpy(s.key + 4, &len, 4); }
        s.kind = kind; s.ttl = ttl; s.val = val;
    }
    // insert or overwrite ("last writer wins", like assigning into a JS object); 2-choice cuckoo
    bool put(uint32_t ns, const uint8_t* k, uint32_t len, uint8_t kind, uint32_t ttl, uint32_t val) {
        uint32_t h2;
        const uint32_t h = hash_key2(ns, k, len, &h2);
        const uint32_t i1 = slot1_of(h, mask), i2 = slot2_of(h, h2, mask);
        if (h2s.size() != (size_t)mask + 1) h2s.assign((size_t)mask + 1, 0);
        for (uint32_t i : { i1, i2 }) {
            Slot& s = z->slots[i];
            if (s.kind != K_EMPTY && s.hash == h && key_eq(s, ns, k, len)) { s.kind = kind; s.ttl = ttl; s.val = val; return false; }
        }
        Slot cur; fill(cur, h, ns, k, len, kind, ttl, val);
        uint32_t cur_h2 = h2;
        uint32_t pos = z->slots[i1].kind == K_EMPTY ? i1 : (z->slots[i2].kind == K_EMPTY ? i2 : i1);
        for (int kick = 0; kick < 2000; kick++) {
            Slot& s = z->slots[pos];
            if (s.kind == K_EMPTY) { s = cur; h2s[pos] = cur_h2; return true; }
            if (getenv("BB_DEBUG") && kick >= 1990) fprintf(stderr, "  kick %d pos=%u resident h=%08x klen=%u ns=%u key=%.*s\n", kick, pos, s.hash, s.klen, s.ns, (int)(s.klen < 49 ? s.klen : 8), (const char*)s.key);
            Slot ev = s; s = cur; cur = ev;                          // evict the resident, move it to its other slot
            const uint32_t ev_h2 = h2s[pos]; h2s[pos] = cur_h2; cur_h2 = ev_h2;
            const uint32_t a = slot1_of(cur.hash, mask), b = slot2_of(cur.hash, cur_h2, mask);
            pos = pos == a ? b : a;
        }
        if (getenv("BB_DEBUG")) fprintf(stderr, "cuckoo fail: ns=%u len=%u key=%.*s h=%08x i1=%u i2=%u mask=%u\n", ns, len, (int)len, (const char*)k, h, i1, i2, mask);
        failed = true;
        return true;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
struct bb_zone { bb::ZoneImage img; };

extern "C" bb_zone* bb_zone_build_shard(const char* buf, size_t len, const char* dns_domain, uint32_t nranks,
                                        uint32_t rank, int* err) {
    auto fail = [&](int e) -> bb_zone* { if (err) *err = e; return nullptr; };
    if (err) *err = BB_OK;
    if ((!buf && len) || !dns_domain || nranks == 0 || rank >= nranks) return fail(BB_ERR_ARG);
    Builder B;
    B.dns_domain = dns_domain;
    // ZKCache.isReady() compares with options.domain verbatim (lib/zk.js:55-58) while keys are
    // lower-cased (:84): an upper-case domain is never ready.  We require lower case instead.
    for (char c : B.dns_domain) if (c >= 'A' && c <= 'Z') return fail(BB_ERR_DOMAIN);
    std::string w;
    if (B.dns_domain.empty() || !to_wire(B.dns_domain.data(), B.dns_domain.size(), w) || w.size() + 1 > 244) return fail(BB_ERR_DOMAIN);
    // root node (lib/zk.js:68-76): exists as soon as the cache is built, data null
    size_t dot = B.dns_domain.find('.');
    std::string first = B.dns_domain.substr(0, dot);
    B.add_node(0, first.data(), first.size());
    std::string root_path;                                     // lib/zk.js:225-228
    {
        std::vector<std::string> parts; size_t s = 0;
        for (;;) { size_t d = B.dns_domain.find('.', s); parts.push_back(B.dns_domain.substr(s, d == std::string::npos ? d : d - s)); if (d == std::string::npos) break; s = d + 1; }
        for (size_t i = parts.size(); i-- > 0;) { root_path += "/"; root_path += parts[i]; }
    }
    std::vector<uint8_t> seen_data(1, 0);
    const char* p = buf; const char* end = buf + len;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* a = p; const char* b = nl ? nl : end;
        p = nl ? nl + 1 : end;
        while (a < b && (*a == ' ' || *a == '\t' || *a == '\r')) ++a;
        while (b > a && (b[-1] == ' ' || b[-1] == '\t' || b[-1] == '\r')) --b;
        if (a == b) continue;
        B.J.clear();
        int ent = B.J.parse(a, b);
        if (ent < 0 || B.J.t[ent].type != J_OBJ) return fail(BB_ERR_SNAPSHOT);
        int path = B.J.get(ent, "path");
        if (path < 0 || B.J.t[path].type != J_STR) return fail(BB_ERR_SNAPSHOT);
        const char* ps = B.J.pool.data() + B.J.t[path].a; size_t pl = B.J.t[path].b;
        uint32_t id;
        if (pl == root_path.size() && !memcmp(ps, root_path.data(), pl)) id = 0;
        else {
            if (pl <= root_path.size() + 1 || memcmp(ps, root_path.data(), root_path.size()) || ps[root_path.size()] != '/') continue;
            size_t i = root_path.size() + 1; uint32_t cur = 0; bool skip = false; int found = -1;
            for (;;) {
                size_t j = i; while (j < pl && ps[j] != '/') ++j;
                if (j == i) { skip = true; break; }               // empty component
                found = B.find_child(cur, ps + i, j - i);
                if (j == pl) {                                     // last component: the znode itself
                    if (found < 0) found = (int)B.add_node(cur, ps + i, j - i);
                    break;
                }
                if (found < 0) { skip = true; break; }            // parent not mirrored
                cur = (uint32_t)found; i = j + 1;
                if (i >= pl) { skip = true; break; }              // trailing '/'
            }
            if (skip) continue;
            id = (uint32_t)found;
        }
        if (seen_data.size() < B.nodes.size()) seen_data.resize(B.nodes.size(), 0);
        if (seen_data[id]) return fail(BB_ERR_SNAPSHOT);           // one line per znode
        seen_data[id] = 1;
        int raw = B.J.get(ent, "raw"), data = B.J.get(ent, "data");
        if (raw >= 0 && B.J.t[raw].type == J_STR) {
            // the znode's bytes: JSON.parse them (lib/zk.js:141-148); failure = ignored
            std::string rs(B.J.pool.data() + B.J.t[raw].a, B.J.t[raw].b);
            int v = B.J.parse(rs.data(), rs.data() + rs.size());
            if (v >= 0 && (B.J.t[v].type == J_NULL || B.J.is_obj(v))) B.ingest(id, v);     // :149-155
        } else if (data >= 0) {
            if (B.J.t[data].type == J_NULL || B.J.is_obj(data)) B.ingest(id, data);
        }
    }

    // ---- lay out the table ------------------------------------------------------------
    bb_zone* zone = new bb_zone();
    ZoneImage& Z = zone->img;
    memset(&Z, 0, sizeof Z);
    TableBuilder T;
    for (uint32_t grow = 0;; grow++) {      // cuckoo insertion can (rarely) fail: rebuild one size up
    uint64_t nkeys = 0;
    for (auto& nd : B.nodes) nkeys += 1 + ((nd.flags & NF_REV) ? 1 : 0);
    // 2-choice cuckoo needs a load factor below 0.5: size for <= 0.45 (a shard holds ~1/nranks of the keys)
    uint64_t want = (nkeys * 22 / 10) / nranks + (nranks > 1 ? nkeys / (4 * nranks) : 0) + 64; uint32_t ns = 64;
    while (ns < want) { ns <<= 1; if (ns == 0) { delete zone; return fail(BB_ERR_NOMEM); } }
    ns <<= grow;
    free(Z.slots); Z.slots = nullptr; Z.n_fwd = Z.n_rev = 0;
    Z.nslots = ns;
    Z.slots = (Slot*)aligned_alloc(64, (size_t)ns * sizeof(Slot));
    if (!Z.slots) { delete zone; return fail(BB_ERR_NOMEM); }
    memset(Z.slots, 0, (size_t)ns * sizeof(Slot));
    T = TableBuilder(); T.z = &Z; T.mask = ns - 1; T.nranks = nranks; T.rank = rank;
    T.arena.assign(4, 0);                                     // offset 0 is never a valid record
    Z.n_nodes = B.nodes.size();
    std::string dom, kw, tmp;
    for (uint32_t id = 0; id < B.nodes.size(); id++) {
        const Node& nd = B.nodes[id];
        B.domain_of(id, dom);
        bool dom_ok = to_wire(dom.data(), dom.size(), tmp) && tmp.size() + 1 <= 255;
        std::string dom_wire = tmp;
        // ---- forward key (lib/zk.js:96) -------------------------------------------------
        if (dom_ok && T.mine(NS_FORWARD, (const uint8_t*)dom.data(), (uint32_t)dom.size())) {   // unspellable names are unreachable; other ranks own the rest
            uint32_t val = nd.addr;
            if (nd.kind == K_SERVICE) {
                const SvcInfo& si = B.svcs[nd.extra];
                std::vector<uint32_t> kids;
                for (uint32_t k = nd.first_kid; k; k = B.nodes[k].next_sib) if (B.nodes[k].flags & NF_KIDTYPE) kids.push_back(k);
                if (kids.size() > 65535) { bb_zone_free(zone); return fail(BB_ERR_SNAPSHOT); }
                std::vector<uint8_t> rec;
                SvcHdr h; h.ttl = si.ttl; h.nkids = (uint16_t)kids.size(); h.rec_len = 0;
                bool s_ok = si.has_srvce && si.srvce.size() < 255, p_ok = si.has_proto && si.proto.size() < 255;
                h.srvce_len = s_ok ? (uint8_t)si.srvce.size() : 0xFF;
                h.proto_len = p_ok ? (uint8_t)si.proto.size() : 0xFF;
                rec.insert(rec.end(), (uint8_t*)&h, (uint8_t*)&h + sizeof h);
                if (s_ok) rec.insert(rec.end(), si.srvce.begin(), si.srvce.end());
                if (p_ok) rec.insert(rec.end(), si.proto.begin(), si.proto.end());
                while (rec.size() & 3) rec.push_back(0);
                size_t tab = rec.size();
                rec.resize(tab + 4 * kids.size());
                while (T.arena.size() & 3) T.arena.push_back(0);
                uint32_t base = (uint32_t)T.arena.size();
                for (size_t ki = 0; ki < kids.size(); ki++) {
                    const Node& kn = B.nodes[kids[ki]];
                    while (rec.size() & 3) rec.push_back(0);
                    uint32_t koff = base + (uint32_t)rec.size();
                    memcpy(rec.data() + tab + 4 * ki, &koff, 4);
                    KidRec kr; memset(&kr, 0, sizeof kr);
                    std::vector<uint16_t> pl;
                    bool name_ok = to_wire(B.pool.data() + kn.name_off, kn.name_len, kw) && kw.size() + dom_wire.size() + 1 <= 255;
                    if (!(kn.flags & NF_SUB_OBJ)) kr.flags = KID_BAD_A | KID_BAD_SRV;            // :366-376
                    else if (kn.flags & NF_ADDR_NULL) kr.flags = KID_ADDR_NULL;                    // :378-381
                    else {
                        bool bad = !(kn.flags & NF_ADDR_OK) || !(kn.flags & NF_TTL_OK);
                        bool bad_srv = bad || !name_ok || (kn.flags & NF_PORTS_BAD);
                        if (kn.flags & NF_PORTS_LIST) pl = B.ports[kn.extra];
                        else if (si.port_ok) pl.push_back((uint16_t)si.port);                     // :383-385
                        else bad_srv = true;
                        kr.flags = (bad ? KID_BAD_A : 0) | (bad_srv ? KID_BAD_SRV : 0);
                        if (kn.flags & NF_HAS_TTL) kr.flags |= KID_HAS_RTTL;
                        kr.addr = kn.addr; kr.rttl = kn.ttl;
                        if (bad_srv) pl.clear();
                    }
                    if (!name_ok) kw.clear();
                    kr.wire_len = (uint8_t)kw.size(); kr.nports = (uint8_t)pl.size();
                    rec.insert(rec.end(), (uint8_t*)&kr, (uint8_t*)&kr + sizeof kr);
                    rec.insert(rec.end(), (uint8_t*)pl.data(), (uint8_t*)pl.data() + 2 * pl.size());
                    rec.insert(rec.end(), kw.begin(), kw.end());
                }
                { uint32_t rl = (uint32_t)rec.size(); memcpy(rec.data() + offsetof(SvcHdr, rec_len), &rl, 4); }
                val = T.arena_put(rec.data(), rec.size());
            }
            uint32_t ttl = nd.kind == K_SERVICE ? B.svcs[nd.extra].ttl : nd.ttl;
            if (T.put(NS_FORWARD, (const uint8_t*)dom.data(), (uint32_t)dom.size(), nd.kind, ttl, val)) Z.n_fwd++;
        }
        // ---- reverse key (lib/zk.js:183-188) ---------------------------------------------
        if ((nd.flags & NF_REV) && nd.rev_len <= 253 && T.mine(NS_REVERSE, (const uint8_t*)B.pool.data() + nd.rev_off, nd.rev_len)) {
            uint8_t kind = K_PTR_BAD; uint32_t val = 0;
            if ((nd.flags & NF_TTL_OK) && dom_ok) {          // lib/server.js:123-130
                std::string t; t.push_back((char)(dom_wire.size() + 1)); t += dom_wire; t.push_back(0);
                val = T.arena_put(t.data(), t.size()); kind = K_PTR;
            }
            if (T.put(NS_REVERSE, (const uint8_t*)B.pool.data() + nd.rev_off, nd.rev_len, kind, nd.ttl, val)) Z.n_rev++;
        }
    }
    if (!T.failed) break;
    if (grow > 3) { bb_zone_free(zone); return fail(BB_ERR_NOMEM); }
    }   // grow
    while (T.arena.size() & 15) T.arena.push_back(0);
    Z.arena_len = T.arena.size();
    Z.arena = (uint8_t*)aligned_alloc(64, (Z.arena_len + 63) & ~(uint64_t)63);
    if (!Z.arena) { bb_zone_free(zone); return fail(BB_ERR_NOMEM); }
    memcpy(Z.arena, T.arena.data(), Z.arena_len);
    Z.ready = 1;                                              // the root TreeNode exists (lib/zk.js:55-58)
    return zone;
}

extern "C" bb_zone* bb_zone_build(const char* buf, size_t len, const char* dns_domain, int* err) {
    return bb_zone_build_shard(buf, len, dns_domain, 1, 0, err);
}

extern "C" void bb_zone_free(bb_zone* z) {
    if (!z) return;
    free(z->img.slots); free(z->img.arena);
    delete z;
}

extern "C" uint64_t bb_zone_stat(const bb_zone* z, int what) {
    if (!z) return 0;
    switch (what) {
    case 0: return z->img.n_nodes;
    case 1: return z->img.n_fwd;
    case 2: return z->img.n_rev;
    case 3: return z->img.nslots;
    case 4: return (uint64_t)z->img.nslots * sizeof(bb::Slot) + z->img.arena_len;
    case 5: return z->img.arena_len;
    }
    return 0;
}

// used by engine.cu
extern "C" const bb::ZoneImage* bb_zone_image(const bb_zone* z) { return z ? &z->img : nullptr; }
