// binder_b200 engine: the CUDA resolve kernel for sm_100a and the host side of the C ABI.
//
// One kernel does the whole of binder's per-query path for a batch of raw DNS packets:
//   mname decode                      -> decode()                 (call site lib/server.js:443-446,471)
//   onQuery type dispatch             -> resolve_query()          (lib/server.js:491-506)
//   resolve / resolvePtr              -> resolve_forward/_ptr()   (lib/server.js:67-134,136-429)
//   zkCache.lookup / reverseLookup    -> probe()                  (lib/zk.js:62-67)
//   shuffle                           -> make_perm()/perm_at()    (lib/server.js:40-53)
//   mname encode + respond            -> emit_response()          (lib/server.js:130,286,299,310,398-402,413-414,427)
//   miss hand-off to recursion        -> compacted miss_idx[]     (lib/server.js:110-113,222-225)
//
// Data movement (HBM-bound integer/byte work, no tensor cores):
//   * a CTA owns a tile of 128 consecutive queries; their packed bytes are one contiguous
//     range of the input, staged into shared memory with coalesced 16-byte loads;
//   * each thread parses its packet from shared memory, hashes the normalised name and
//     probes the zone table in HBM (one 64-byte slot = two sectors per host record);
//   * response sizes are scanned in the CTA, tile bases come from a single-pass decoupled
//     look-back across CTAs (so output is packed, in query order, in ONE kernel);
//   * responses are assembled in shared memory and flushed with 16-byte coalesced stores.
#include "zone_image.h"
#include "../../include/binder_b200.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" const bb::ZoneImage* bb_zone_image(const bb_zone* z);
extern "C" int bb_zone_pending(const bb_zone* z, const uint32_t** slots, uint32_t* n_slots, uint64_t* arena_from, int* relaid);
extern "C" void bb_zone_mark_synced(bb_zone* z);
extern "C" uint64_t bb_zone_sync_gen(const bb_zone* z);

namespace bbk {
using namespace bb;

constexpr int T = 128;                    // queries (= threads) per tile
constexpr int S_IN = 8192;                // staged input bytes per tile
constexpr int CAPW = 12288;               // output staging window per flush round
constexpr int MAXRESP = 1232;             // >= the largest response (1200)
constexpr int S_OUT = ((CAPW + 32 + 127) / 128 + 1) * 128;   // whole 128-byte rows (the staging buffer is swizzled per row)
constexpr uint32_t NONE16 = 0xFFFF;

constexpr uint64_t D_FLAG_A = 1ull << 62, D_FLAG_P = 2ull << 62, D_VAL = (1ull << 62) - 1;
constexpr int D_MISS_SHIFT = 40;

enum { ST_ANSWERED = 0, ST_MISS = 1, ST_DROPPED = 2 };
enum { RC_NOERROR = 0, RC_SERVFAIL = 2, RC_NXDOMAIN = 3, RC_NOTIMP = 4, RC_REFUSED = 5 };
enum { QT_A = 1, QT_SOA = 6, QT_PTR = 12, QT_SRV = 33, QT_OPT = 41 };
enum { RK_NONE = 0, RK_HEADER = 1, RK_A1 = 2, RK_PTR = 3, RK_SOA = 4, RK_SVC_A = 5, RK_SVC_SRV = 6 };

struct Params {
    const uint8_t* pkts; const uint32_t* pkt_off; uint32_t n;
    uint64_t seed; uint32_t qidx_base;
    uint8_t* out; uint32_t out_cap; uint32_t* out_off; uint16_t* out_len; uint8_t* status; uint32_t* miss_idx; uint32_t* totals;
    const Slot* table; uint32_t mask; const uint8_t* arena; int ready;
    const EngineConst* eng;
    unsigned long long* desc; uint32_t* counter; uint32_t ntiles, ntiles_cap;   // desc[ntiles_cap] = arrival cursor
    uint32_t epoch;          // launch number: marks totals[2] (overflow) / totals[3] (done) of THIS launch
    unsigned long long* stage_log;   // optional [ntiles][8] globaltimer stamps (bb_engine_set_stage_log)
    const uint32_t* n_dev;           // when set, the batch size is read from device memory (routed batches)
    const uint32_t* qidx_map;        // when set, query i's shuffle index is qidx_map[i] (routed batches)
    uint32_t route, nranks, rank;    // route != 0: compute the owner rank of each query instead of probing
    uint32_t suffix_len, soa_len, recursion;   // copies of EngineConst scalars (constant bank instead of a global load)
    uint32_t tcp;            // the batch arrived over TCP: no 512-byte / EDNS size limit (RFC 1035 4.2.2)
    uint32_t* qidx_out;      // multi-region: each result's ingress index is also written here (host result mirrors)
    const uint32_t* err_in;  // multi-region: the shard's wait-timeout word, copied into totals[6]
    uint8_t* bounce;         // zero-copy results: device buffer (same offsets as `out`) that direct-emit tiles write to
                             // before copying their range to the host buffer with coalesced stores
    // multi-region launch (grid.y = regions): every per-batch pointer advances by its stride per region
    uint32_t regions;
    size_t in_stride, out_stride, off_stride, len_stride, status_stride, miss_stride, totals_stride, desc_stride;
};

// per-thread state carried from the sizing pass to the emit pass
struct Res {
    const uint8_t* p;        // packet bytes (shared memory, or global when the tile did not fit)
    uint64_t lenmask;        // bit i set: QNAME byte i is a label-length byte (names <= 64 bytes)
    uint32_t sp;             // shared-memory address of the packet (0 when not staged)
    uint32_t qn_len;         // QNAME wire length incl. terminator
    uint32_t ttl, val;
    uint64_t perm;           // shuffled child order, 4 bits each (nk <= 16)
    uint16_t rlen, maxsz, qtype, adv;
    uint16_t d_off, d_end;   // domain part [d_off, d_end) in QNAME wire coordinates
    uint16_t ptr_tgt;        // label boundary the owner's compression pointer targets, or NONE16
    uint16_t lastlen;        // position of the domain's last length byte
    uint16_t keep_ans, keep_add, n_walk, nk;
    uint8_t status, rk, rcode, tc, opcode, rd, edns, trunc;
    uint8_t owner;           // route mode: rank that owns this query's lookup key
};

__device__ __forceinline__ uint32_t lower8(uint32_t c) { return (c - 'A' < 26u) ? c + 32 : c; }
__device__ __forceinline__ uint32_t be16(const uint8_t* p) { return (uint32_t)p[0] << 8 | p[1]; }
__device__ __forceinline__ uint32_t ld32(const uint8_t* p) { return *(const uint32_t*)p; }    // 4-byte aligned
__device__ __forceinline__ uint32_t ld16a(const uint8_t* p) { return *(const uint16_t*)p; }   // 2-byte aligned

// byte-fed murmur (same result as bb::hash_key over the materialised key)
struct KeyHash {
    uint32_t h, g, acc, n;
    __device__ void init(uint32_t ns) { h = hash_init(ns); g = hash2_init(ns); acc = 0; n = 0; }
    __device__ void feed(uint32_t c) {
        acc |= c << (8 * (n & 3)); ++n;
        if ((n & 3) == 0) { h = hash_word(h, acc); g = hash2_word(g, acc); acc = 0; }
    }
    // -> primary hash; h2 = second hash (second cuckoo slot)
    __device__ uint32_t finish(uint32_t& h2) {
        if (n & 3) { h = hash_word(h, acc); g = hash2_word(g, acc); }
        h2 = hash2_finish(g, n);
        return hash_finish(h, n);
    }
};

// ---- mname decode (DESIGN.md "Wire spec: decode") ---------------------------------------
__device__ bool decode(const uint8_t* p, uint32_t len, Res& r) {
    if (len < 12) return false;
    if (p[2] & 0x80) return false;
    r.opcode = (p[2] >> 3) & 0xF; r.rd = p[2] & 1;
    uint32_t qd = be16(p + 4), an = be16(p + 6), ns = be16(p + 8), ar = be16(p + 10);
    if (qd != 1 || an != 0 || ns != 0 || ar > 1) return false;
    uint32_t pos = 12;
    uint64_t lm = 0;
    for (;;) {
        if (pos >= len) return false;
        uint32_t c = p[pos];
        if (c == 0) { ++pos; break; }
        if (c > 63 || pos + 1 + c > len) return false;
        if (pos - 12 < 64) lm |= 1ull << (pos - 12);
        pos += 1 + c;
        if (pos - 12 + 1 > 255) return false;
    }
    r.lenmask = lm;
    r.qn_len = pos - 12;
    if (pos + 4 > len) return false;
    r.qtype = (uint16_t)be16(p + pos);
    if (be16(p + pos + 2) != 1) return false;
    pos += 4;
    r.edns = 0; r.adv = 0;
    if (ar == 1) {
        if (pos + 11 > len || p[pos] != 0 || be16(p + pos + 1) != QT_OPT) return false;
        r.adv = (uint16_t)be16(p + pos + 3);
        if (pos + 11 + be16(p + pos + 9) > len) return false;
        r.edns = 1;
    }
    return true;
}

// ---- zkCache.lookup / reverseLookup ------------------------------------------------------
// The key is produced twice (hash, then compare) by the same generator so that nothing is
// materialised.  Forward keys: the domain part in dotted lower case.  Reverse keys: the
// labels before "in-addr.arpa", reversed, joined by '.'.
struct FwdKey {
    const uint8_t* nm; uint32_t d_off, d_end;
    uint32_t pos, nlp;
    __device__ uint32_t length() const { return d_end - d_off - 1; }
    __device__ void start() { pos = d_off + 1; nlp = d_off + 1 + nm[d_off]; }
    __device__ uint32_t next() {
        uint32_t c;
        if (pos == nlp) { c = '.'; nlp = pos + 1 + nm[pos]; } else c = lower8(nm[pos]);
        ++pos; return c;
    }
};
struct RevKey {
    const uint8_t* nm; uint32_t nlab;       // labels before in-addr.arpa
    uint32_t len_;
    int k; uint32_t pos, rem; bool dot;
    __device__ uint32_t label_pos(int idx) const { uint32_t q = 0; for (int i = 0; i < idx; i++) q += 1 + nm[q]; return q; }
    __device__ void measure() { len_ = 0; uint32_t q = 0; for (uint32_t i = 0; i < nlab; i++) { len_ += nm[q] + (i ? 1 : 0); q += 1 + nm[q]; } }
    __device__ uint32_t length() const { return len_; }
    __device__ void start() { k = (int)nlab - 1; dot = false; if (k >= 0) { pos = label_pos(k); rem = nm[pos]; ++pos; } }
    __device__ uint32_t next() {
        if (dot) { dot = false; return '.'; }
        uint32_t c = nm[pos++]; --rem;
        if (rem == 0 && k > 0) { --k; pos = label_pos(k); rem = nm[pos]; ++pos; dot = true; }
        return c;
    }
};

template <class KG>
__device__ bool probe(const Params& P, Res& r, uint32_t ns, KG& kg, uint32_t& kind, uint32_t& ttl, uint32_t& val) {
    uint32_t klen = kg.length();
    KeyHash kh; kh.init(ns);
    kg.start();
    for (uint32_t i = 0; i < klen; i++) kh.feed(kg.next());
    uint32_t h2;
    uint32_t h = kh.finish(h2);
    if (P.route) { r.owner = (uint8_t)owner_of(h, P.nranks); return false; }   // sharding: who would answer
    // 2-choice cuckoo: the key is in slot1_of(h) or slot2_of(h) or nowhere
    const uint32_t cand[2] = { slot1_of(h, P.mask), slot2_of(h, h2, P.mask) };
    for (int c = 0; c < 2; c++) {
        const Slot* s = P.table + cand[c];
        uint4 hd = __ldg((const uint4*)s);                  // hash | klen,kind,ns,flags | ttl | val
        uint32_t sk = (hd.y >> 8) & 0xFF;
        if (sk == K_EMPTY) continue;
        if (hd.x == h && ((hd.y >> 16) & 1) == ns) {
            uint32_t sl = hd.y & 0xFF;
            const uint8_t* kb = nullptr;
            if (sl == KLEN_OVERFLOW) {
                uint32_t off = __ldg((const uint32_t*)s->key), l = __ldg((const uint32_t*)(s->key + 4));
                if (l == klen) kb = P.arena + off;
            } else if (sl == klen) kb = s->key;
            if (kb) {
                kg.start();
                bool eq = true;
                for (uint32_t j = 0; j < klen; j++) if (__ldg(kb + j) != kg.next()) { eq = false; break; }
                if (eq) { kind = sk; ttl = hd.z; val = hd.w; return true; }
            }
        }
    }
    return false;
}

// ---- shuffle (lib/server.js:40-53) --------------------------------------------------------
__device__ uint64_t make_perm(uint32_t n, uint64_t seed, uint32_t qidx) {     // n <= 16
    uint64_t perm = 0xFEDCBA9876543210ull;
    for (uint32_t i = n; i-- > 1;) {
        uint32_t j = shuffle_rand(seed, qidx, i);
        uint64_t x = ((perm >> (4 * i)) ^ (perm >> (4 * j))) & 15;
        perm ^= (x << (4 * i)) | (x << (4 * j));
    }
    return perm;
}
// element that ends up at position `p`, for any n: undo the swaps in reverse order
__device__ uint32_t perm_at_slow(uint32_t p, uint32_t n, uint64_t seed, uint32_t qidx) {
    uint32_t pos = p;
    for (uint32_t i = 1; i < n; i++) {
        uint32_t j = shuffle_rand(seed, qidx, i);
        if (pos == i) pos = j; else if (pos == j) pos = i;
    }
    return pos;
}
__device__ __forceinline__ uint32_t perm_at(const Res& r, uint32_t t, uint64_t seed, uint32_t qidx) {
    return r.nk <= 16 ? (uint32_t)(r.perm >> (4 * t)) & 15 : perm_at_slow(t, r.nk, seed, qidx);
}

struct SvcView {
    const uint8_t* base; const uint8_t* arena; const uint32_t* kid_off;
    __device__ void open(const uint8_t* arena_, uint32_t off) {
        arena = arena_; base = arena_ + off;
        const SvcHdr* h = (const SvcHdr*)base;
        uint32_t sl = h->srvce_len == 0xFF ? 0 : h->srvce_len, pl = h->proto_len == 0xFF ? 0 : h->proto_len;
        kid_off = (const uint32_t*)(base + ((sizeof(SvcHdr) + sl + pl + 3) & ~3u));
    }
    __device__ const SvcHdr* hdr() const { return (const SvcHdr*)base; }
    __device__ const KidRec* kid(uint32_t i) const { return (const KidRec*)(arena + kid_off[i]); }
};

// owner-name sizes for this query's domain part (DESIGN.md "Wire spec: compression")
__device__ __forceinline__ uint32_t dom_owner_len(const Res& r) {
    return r.ptr_tgt != NONE16 ? (uint32_t)(r.ptr_tgt - r.d_off) + 2 : (uint32_t)(r.d_end - r.d_off) + 1;
}
__device__ __forceinline__ uint32_t dom_wire_len(const Res& r) { return (uint32_t)(r.d_end - r.d_off) + 1; }

// Sizing pass over a service's children in shuffled order (lib/server.js:361-416).
__device__ __forceinline__ unsigned long long gtime_early() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STAMP_SVC(k) do { if (P.stage_log && threadIdx.x == 0) P.stage_log[(size_t)blockIdx.x * 16 + (k)] = gtime_early(); } while (0)
__device__ void size_service(const Params& P, Res& r, uint32_t qidx, bool srv, uint32_t fixed) {
    STAMP_SVC(11);
    SvcView sv; sv.open(P.arena, r.val);
    {   // The record (header, child offsets, children) is contiguous: touch all of its cache lines now,
        // with independent loads, so that the dependent walks below (and in the emit pass) hit in
        // L1/L2 instead of paying a DRAM round trip per child.
        const uint32_t rl = sv.hdr()->rec_len;
        const uint8_t* b0 = (const uint8_t*)((uintptr_t)sv.base & ~(uintptr_t)127);
        const uint8_t* e0 = sv.base + rl;
        for (const uint8_t* q = b0 + 128; q < e0; q += 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(q));
    }
    uint32_t nk = sv.hdr()->nkids;
    r.nk = (uint16_t)nk;
    STAMP_SVC(12);
    r.perm = nk <= 16 ? make_perm(nk, P.seed, qidx) : 0;
    STAMP_SVC(13);
    const uint32_t dol = dom_owner_len(r), dwl = dom_wire_len(r);
    uint32_t ans_b = 0, add_b = 0, n_ans = 0, n_add = 0, n_walk = nk;
    const uint8_t badbit = srv ? KID_BAD_SRV : KID_BAD_A;
    for (uint32_t t = 0; t < nk; t++) {
        const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
        uint32_t fl = k->flags;
        if (fl & badbit) { r.rcode = RC_SERVFAIL; n_walk = t; break; }       // :366-376
        if (fl & KID_ADDR_NULL) continue;                                     // :378-381
        if (srv) {
            ans_b += (uint32_t)k->nports * (18 + k->wire_len + dwl); n_ans += k->nports;
            add_b += k->wire_len + dol + 14; n_add++;
        } else { ans_b += dol + 14; n_ans++; }
    }
    r.n_walk = (uint16_t)n_walk;
    STAMP_SVC(14);
    if (fixed + ans_b + add_b <= r.maxsz) { r.keep_ans = (uint16_t)n_ans; r.keep_add = (uint16_t)n_add; r.rlen = (uint16_t)(fixed + ans_b + add_b); return; }
    // truncation: keep the longest prefix of [answers..., additionals...] that fits
    r.tc = 1;
    uint32_t total = fixed, ka = 0, kd = 0; bool full = false;
    for (uint32_t t = 0; t < n_walk && !full; t++) {
        const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
        if (k->flags & KID_ADDR_NULL) continue;
        uint32_t each = srv ? 18 + k->wire_len + dwl : dol + 14, cnt = srv ? k->nports : 1;
        for (uint32_t c = 0; c < cnt; c++) { if (total + each > r.maxsz) { full = true; break; } total += each; ++ka; }
    }
    if (!full && srv) for (uint32_t t = 0; t < n_walk; t++) {
        const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
        if (k->flags & KID_ADDR_NULL) continue;
        uint32_t each = k->wire_len + dol + 14;
        if (total + each > r.maxsz) break;
        total += each; ++kd;
    }
    r.keep_ans = (uint16_t)ka; r.keep_add = (uint16_t)kd; r.rlen = (uint16_t)total;
}

// one RR that either fits or is dropped (TC)
__device__ __forceinline__ void size_single(Res& r, uint32_t fixed, uint32_t rr) {
    if (fixed + rr <= r.maxsz) { r.rlen = (uint16_t)(fixed + rr); r.keep_ans = 1; }
    else { r.rlen = (uint16_t)fixed; r.keep_ans = 0; r.tc = 1; }
}

// Recursion.resolve()'s quick rejects (lib/recursion.js:329-344), for a miss that would otherwise be handed
// to the host: would it be forwarded anywhere?  nm = QNAME wire bytes — query.name() in its original case,
// SRV prefix included — W = its length without the terminator.  In the dotted string every label boundary
// is a '.', so the string operations map one to one onto wire positions:
//   domain.indexOf(dnsDomain, domain.length - dnsDomain.length) === -1            -> not ours   (:330-333)
//   p = domain minus the suffix and the one character before it; dc = p after its last '.';
//   self.dcs[dc] === undefined (or every upstream of dc is this host)              -> nowhere to ask (:338-343,377-379)
// Kept out of line: it runs for misses only and must not cost the hit path registers.
__device__ __noinline__ bool recursion_forwardable(const EngineConst* E, const uint8_t* nm, uint32_t W) {
    const uint32_t L = E->rf_dom_len;
    if (W < 1 + L) return false;                       // name shorter than the suffix
    const uint32_t s0 = W - L;                         // wire index where the suffix must start
    if (s0 < 2) return false;                          // nothing before it: dc = ''
    const uint32_t cw = s0 - 1;                        // the character substring() drops (normally the '.')
    uint32_t nlp = 1u + nm[0], last_b = 0;             // next length byte; last boundary before cw
    bool ok = true;
    for (uint32_t w = 1; w < W; w++) {
        const bool boundary = w == nlp;
        if (boundary) { nlp = w + 1u + nm[w]; if (w < cw) last_b = w; }
        if (w >= s0) {
            const uint32_t e = E->rf_dom[w - s0];
            ok &= boundary ? e == '.' : (e != '.' && e == nm[w]);
        }
    }
    if (!ok) return false;
    const uint32_t d0 = last_b + 1, dl = cw - d0;      // dc = dotted[d0 .. cw)
    if (dl == 0 || dl > 63) return false;
    for (uint32_t k = 0; k < E->rf_ndc; k++) {
        if (E->rf_dc_len[k] != dl) continue;
        bool eq = true;
        for (uint32_t i = 0; i < dl; i++) eq &= E->rf_dc[k][i] == nm[d0 + i];
        if (eq) return true;
    }
    return false;
}

// What resolve() does once zk.lookup() has answered (lib/server.js:219-424); shared by the
// generic and the word-wise front ends.
__device__ void finish_forward(const Params& P, Res& r, uint32_t qidx, uint32_t fixed, bool srv, bool hit,
                               uint32_t kind, uint32_t ttl, uint32_t val, uint32_t l0, uint32_t l1) {
    const uint8_t* nm = r.p + 12;
    const EngineConst* E = P.eng;
    if (!hit) {                                                               // :219-247
        if (P.recursion && r.rd) {
            // pre-filter: a miss recursion.js would refuse without asking anyone is refused here (same bytes)
            if (P.recursion == 2 && !recursion_forwardable(E, nm, r.qn_len - 1)) { r.rcode = RC_REFUSED; return; }
            r.status = ST_MISS; r.rk = RK_NONE; r.rlen = 0; return;
        }
        r.rcode = RC_REFUSED; return;
    }
    r.ttl = ttl; r.val = val;
    if (kind == K_INVALID) { r.rcode = RC_SERVFAIL; return; }                 // :251-260
    if (srv && kind != K_SERVICE) {                                           // :276-292 NODATA + SOA
        r.rcode = RC_NOERROR; r.rk = RK_SOA;
        uint32_t rr = dom_owner_len(r) + 10 + E->soa_len + 20;
        if (fixed + rr <= r.maxsz) { r.rlen = (uint16_t)(fixed + rr); r.keep_ans = 1; }
        else { r.keep_ans = 0; r.tc = 1; }
        return;
    }
    if (kind == K_ADDR) { r.rcode = RC_NOERROR; r.rk = RK_A1; size_single(r, fixed, dom_owner_len(r) + 14); return; }
    if (kind == K_ADDR_BAD) { r.rcode = RC_SERVFAIL; return; }                // contract
    if (kind == K_UNKNOWN) { r.rcode = RC_NOTIMP; return; }                   // :419-424 + :346-350
    // K_SERVICE (:313-417)
    SvcView sv; sv.open(P.arena, val);
    const SvcHdr* h = sv.hdr();
    r.ttl = h->ttl;
    if (srv) {
        const uint8_t* sb = sv.base + sizeof(SvcHdr);
        bool match = h->srvce_len == l0 && h->proto_len == l1;
        for (uint32_t i = 0; match && i < l0; i++) if (sb[i] != nm[1 + i]) match = false;
        for (uint32_t i = 0; match && i < l1; i++) if (sb[l0 + i] != nm[2 + l0 + i]) match = false;
        if (!match) { r.rcode = RC_NXDOMAIN; return; }                        // :334-345
    }
    r.rcode = RC_NOERROR;                                                     // :351
    r.rk = srv ? RK_SVC_SRV : RK_SVC_A;
    size_service(P, r, qidx, srv, fixed);
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// per-stage stamps of one tile, the batched analogue of query._stamp() (lib/server.js:479-483)
constexpr int NSTAGE = 16;
#define STAMP(k) do { if (P.stage_log && threadIdx.x == 0) P.stage_log[(size_t)blockIdx.x * NSTAGE + (k)] = gtime(); } while (0)

// ---- word-wise front end of resolve() -----------------------------------------------------
// Same decisions as resolve_forward() below, four name bytes per step, for the common case:
// packet staged in shared memory, QNAME <= 64 wire bytes, lookup key <= 48 bytes (inline slot
// keys).  Anything else returns false and takes the generic path.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// unaligned 32-bit load from shared memory (the staging buffers carry read slack)
__device__ __forceinline__ uint32_t ldsu32(uint32_t a) {
    const uint32_t b = a & ~3u;
    return __funnelshift_r(lds32(b), lds32(b + 4), (a & 3u) * 8);
}
// v << n with PTX semantics: any n > 31 (including a wrapped-around negative) gives 0
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
// 0x80 in every byte of v that is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v) { return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu); }
// 0x80 in every byte of x7 (7-bit bytes) that is >= k
__device__ __forceinline__ uint32_t ge7(uint32_t x7, uint32_t k) { return (x7 + (0x80u - k) * 0x01010101u) & 0x80808080u; }
// 0x80 in every byte that is 'A'..'Z'
__device__ __forceinline__ uint32_t upper_bytes(uint32_t x) {
    const uint32_t x7 = x & 0x7F7F7F7Fu;
    return ge7(x7, 0x41) & ~ge7(x7, 0x5B) & ~x;
}

// Decode of a packet staged in shared memory, word-wise (same acceptance as decode()).
__device__ bool decode_staged(uint32_t sp, uint32_t len, Res& r) {
    if (len < 17) return false;                                               // header + root name + type/class at least
    const uint32_t hb = sp & ~3u, hs = (sp & 3u) * 8;
    const uint32_t t0 = lds32(hb), t1 = lds32(hb + 4), t2 = lds32(hb + 8), t3 = lds32(hb + 12);
    const uint32_t w0 = __funnelshift_r(t0, t1, hs), w1 = __funnelshift_r(t1, t2, hs), w2 = __funnelshift_r(t2, t3, hs);
    const uint32_t fl = (w0 >> 16) & 0xFF;                                    // byte 2: QR opcode AA TC RD
    if (fl & 0x80) return false;
    r.opcode = (fl >> 3) & 0xF; r.rd = fl & 1;
    if (w1 != 0x00000100u) return false;                                      // QDCOUNT=1, ANCOUNT=0
    if (w2 != 0u && w2 != 0x01000000u) return false;                          // NSCOUNT=0, ARCOUNT<=1
    const uint32_t nm = sp + 12, lim = len - 12;                              // name bytes available
    // label hop: one dependent shared-memory byte per label; validity is accumulated, not branched on
    uint32_t pos = 0, lo = 0, hi = 0, bad = 0, c = lds8(nm);
#pragma unroll 1
    while (c != 0) {
        bad |= c > 63;                                                        // pointers / extended label types
        lo |= shl_clamp(1u, pos); hi |= shl_clamp(1u, pos - 32u);             // positions >= 64 fall off (shl.b32 clamps its count)
        pos += 1 + c;
        if (pos >= lim || pos > 254) { bad = 1; break; }
        c = lds8(nm + pos);
    }
    if (bad) return false;
    if (pos + 1 + 4 > lim) return false;
    r.qn_len = pos + 1;
    r.lenmask = (uint64_t)lo | ((uint64_t)hi << 32);
    const uint32_t tc = ldsu32(nm + pos + 1);                                 // QTYPE, QCLASS (big-endian)
    r.qtype = (uint16_t)(((tc & 0xFF) << 8) | ((tc >> 8) & 0xFF));
    if ((tc >> 16) != 0x0100u) return false;                                  // class IN
    r.edns = 0; r.adv = 0;
    if (w2) {
        const uint32_t q = nm + pos + 5;                                      // the one additional RR
        if (pos + 5 + 11 > lim) return false;
        const uint32_t a = ldsu32(q), b = ldsu32(q + 4), c2 = ldsu32(q + 8);
        if ((a & 0xFFFFFF) != 0x290000u) return false;                        // root owner, TYPE 41
        r.adv = (uint16_t)(((a >> 24) << 8) | (b & 0xFF));
        const uint32_t rdlen = (((c2 >> 8) & 0xFF) << 8) | ((c2 >> 16) & 0xFF);
        if (pos + 5 + 11 + rdlen > lim) return false;
        r.edns = 1;
    }
    return true;
}

// 0x80 in every byte of the lower-cased dotted word `lo` that is NOT in [a-z0-9_-] ('.' counts as bad:
// the caller masks out the label-boundary positions)
__device__ __forceinline__ uint32_t bad_chars(uint32_t lo) {
    const uint32_t y7 = lo & 0x7F7F7F7Fu;
    const uint32_t ok = ((ge7(y7, 0x61) & ~ge7(y7, 0x7B)) | (ge7(y7, 0x30) & ~ge7(y7, 0x3A)) |
                         zero_bytes(lo ^ 0x2D2D2D2Du) | zero_bytes(lo ^ 0x5F5F5F5Fu)) & ~lo;
    return ~ok & 0x80808080u;
}

__device__ bool fast_forward(const Params& P, Res& r, uint32_t s_sfx, uint32_t qidx, uint32_t fixed) {
    const EngineConst* E = P.eng;
    if (!P.ready && !P.route) return false;            // not-ready engines: exact ordering of refusals lives in the generic path
    const uint32_t nm = r.sp + 12;
    const bool srv = r.qtype == QT_SRV;
    const uint32_t d_end = r.qn_len - 1;
    uint32_t d_off = 0, l0 = 0, l1 = 0;
    bool refuse = false;
    if (srv) {                                                                // :141-154
        l0 = lds8(nm);
        if (l0 == 0) { r.rcode = RC_REFUSED; return true; }
        const uint32_t p1 = 1 + l0; l1 = lds8(nm + p1);
        if (l1 == 0) { r.rcode = RC_REFUSED; return true; }
        for (uint32_t i = 1; i <= l0; i++) { const uint32_t c = lds8(nm + i); refuse |= (i == 1) ? (c != '_') : (c == '_' || c == '.'); }
        for (uint32_t i = 1; i <= l1; i++) { const uint32_t c = lds8(nm + p1 + i); refuse |= (i == 1) ? (c != '_') : (c == '_' || c == '.'); }
        d_off = p1 + 1 + l1;
        if (lds8(nm + d_off) == 0) { r.rcode = RC_REFUSED; return true; }
    }
    if (d_end <= d_off + 1) { r.rcode = RC_REFUSED; return true; }            // root name
    const uint32_t dl = d_end - d_off - 1;
    if (dl > KEY_INLINE_MAX) return false;
    // suffix gate (:157-166), case-sensitive, on the raw wire bytes: the domain's last sl bytes must be
    // dnsDomain's wire labels and start at a label boundary ('.' + dnsDomain in the dotted view)
    const uint32_t sl = P.suffix_len;
    if (dl < sl) { r.rcode = RC_REFUSED; return true; }       // (a truncated SRV domain is shorter still)
    {
        const uint32_t t0 = d_end - sl;                                       // where the suffix's first length byte must sit
        uint32_t bad = ((r.lenmask >> t0) & 1ull) ? 0u : 1u;
        const uint32_t nw = (sl + 3) >> 2;
        for (uint32_t j = 0; j < nw; j++) {                                   // words right-aligned to the end of the name
            const uint32_t x = ldsu32(nm + d_end - 4 * (j + 1));
            const uint32_t e = lds32(s_sfx + 256 - 4 * (j + 1));
            const uint32_t rem = sl - 4 * j;                                  // bytes of this word that belong to the suffix
            const uint32_t cm = rem >= 4 ? 0xFFFFFFFFu : (0xFFFFFFFFu << (8 * (4 - rem)));
            bad |= (x ^ e) & cm;
        }
        if (bad) { if (srv) return false; r.rcode = RC_REFUSED; return true; }   // SRV: the regex group may stop at a line terminator (:141) -> generic path
    }
    // normalise (dotted view, toLowerCase :207) + hash, four bytes per step
    const uint64_t lm = r.lenmask >> (d_off + 1);
    const uint32_t nwords = (dl + 3) >> 2;
    const uint32_t tailm = (dl & 3) ? ((1u << (8 * (dl & 3))) - 1) : 0xFFFFFFFFu;
    uint32_t kw[12];
    uint32_t h = hash_init(NS_FORWARD), g = hash2_init(NS_FORWARD);
    uint32_t upw = 0, upi = 0;
    // consecutive unaligned words share their aligned halves: one LDS per word, not two
    const uint32_t ka = nm + d_off + 1, kb = ka & ~3u, ksh = (ka & 3u) * 8;
    uint32_t wprev = lds32(kb);
#pragma unroll
    for (int i = 0; i < 12; i++) {
        kw[i] = 0;
        if ((uint32_t)i < nwords) {
            const uint32_t wnext = lds32(kb + 4 * (i + 1));
            const uint32_t x = __funnelshift_r(wprev, wnext, ksh);
            wprev = wnext;
            const uint32_t bits = (uint32_t)(lm >> (4 * i)) & 0xFu;
            const uint32_t m8 = ((bits * 0x00204081u) & 0x01010101u) * 0xFFu;          // label-boundary positions
            uint32_t xd = (x & ~m8) | (0x2E2E2E2Eu & m8);
            if ((uint32_t)i == nwords - 1) xd &= tailm;
            const uint32_t up = upper_bytes(xd);
            if (up) { upw = up; upi = i; }
            const uint32_t lo = xd | (up >> 2);
            kw[i] = lo;
            h = hash_word(h, lo); g = hash2_word(g, lo);
        }
    }
    h = hash_finish(h, dl);
    const uint32_t h2 = hash2_finish(g, dl);
    STAMP(4);
    if (refuse) { r.rcode = RC_REFUSED; return true; }
    if (P.route) { r.owner = (uint8_t)owner_of(h, P.nranks); return true; }   // sharding: who would answer
    r.d_off = (uint16_t)d_off; r.d_end = (uint16_t)d_end; r.trunc = 0; r.lastlen = (uint16_t)d_off;
    if (!upw) r.ptr_tgt = (uint16_t)d_off;
    else {
        const uint32_t pu = d_off + 1 + 4 * upi + ((31 - __clz(upw)) >> 3);  // wire position of the last upper-case byte
        const uint64_t m = pu + 1 < 64 ? (r.lenmask >> (pu + 1)) : 0ull;
        r.ptr_tgt = m ? (uint16_t)(pu + 1 + (__ffsll((long long)m) - 1)) : (uint16_t)NONE16;
    }
    // zk.lookup(domain): one 64-byte slot per probe, compared as words.  The header compare also
    // carries the key's dot count: a query with a '.' inside a label has fewer label boundaries than
    // any key that spells the same, so it can never match here.
    const uint32_t ndots = (uint32_t)__popcll(r.lenmask >> (d_off + 1));
    const uint32_t want = dl | ((NS_FORWARD | (ndots << 1)) << 16);
    uint32_t kind = 0, ttl = 0, val = 0;
    bool hit = false, clean = false;
    {
        // 2-choice cuckoo: both candidate slots are fetched together — one DRAM round trip per lookup,
        // hit or miss, for every lane of the warp
        const uint4* sa = (const uint4*)(P.table + slot1_of(h, P.mask));
        const uint4* sb = (const uint4*)(P.table + slot2_of(h, h2, P.mask));
        const uint4 a0 = __ldg(sa), a1 = __ldg(sa + 1), a2 = __ldg(sa + 2), a3 = __ldg(sa + 3);
        const uint4 b0 = __ldg(sb), b1 = __ldg(sb + 1), b2 = __ldg(sb + 2), b3 = __ldg(sb + 3);
        const uint32_t da = (a0.x ^ h) | ((a0.y & 0x00FF00FFu) ^ want) |
                            (kw[0] ^ a1.x) | (kw[1] ^ a1.y) | (kw[2] ^ a1.z) | (kw[3] ^ a1.w) |
                            (kw[4] ^ a2.x) | (kw[5] ^ a2.y) | (kw[6] ^ a2.z) | (kw[7] ^ a2.w) |
                            (kw[8] ^ a3.x) | (kw[9] ^ a3.y) | (kw[10] ^ a3.z) | (kw[11] ^ a3.w);
        const uint32_t db = (b0.x ^ h) | ((b0.y & 0x00FF00FFu) ^ want) |
                            (kw[0] ^ b1.x) | (kw[1] ^ b1.y) | (kw[2] ^ b1.z) | (kw[3] ^ b1.w) |
                            (kw[4] ^ b2.x) | (kw[5] ^ b2.y) | (kw[6] ^ b2.z) | (kw[7] ^ b2.w) |
                            (kw[8] ^ b3.x) | (kw[9] ^ b3.y) | (kw[10] ^ b3.z) | (kw[11] ^ b3.w);
        // an empty slot has kind 0 and klen 0, so it can never equal `want` (dl >= 1)
        if (da == 0) { hit = true; kind = (a0.y >> 8) & 0xFF; ttl = a0.z; val = a0.w; clean = (a0.y >> 24) & SLOT_KEY_CLEAN; }
        else if (db == 0) { hit = true; kind = (b0.y >> 8) & 0xFF; ttl = b0.z; val = b0.w; clean = (b0.y >> 24) & SLOT_KEY_CLEAN; }
    }
    STAMP(5);
    if (!(hit && clean)) {
        // Not a clean hit: classify the name the way resolve() does before its lookup — a '.' inside a
        // label, a character outside [a-z0-9_.-] (:208-215) -> REFUSED; an SRV name with a
        // line terminator goes to the generic path (its regex group stops there, :141).
        uint32_t bad = 0, nlc = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            if ((uint32_t)i < nwords) {
                const uint32_t bits = (uint32_t)(lm >> (4 * i)) & 0xFu;
                const uint32_t m8 = ((bits * 0x00204081u) & 0x01010101u) * 0xFFu;
                const uint32_t tm = ((uint32_t)i == nwords - 1) ? tailm : 0xFFFFFFFFu;
                bad |= bad_chars(kw[i]) & ~m8 & tm;
                nlc |= (zero_bytes(kw[i] ^ 0x0A0A0A0Au) | zero_bytes(kw[i] ^ 0x0D0D0D0Du)) & ~m8 & tm;
            }
        }
        if (srv && nlc) return false;
        if (bad) { r.rcode = RC_REFUSED; return true; }
    }
    finish_forward(P, r, qidx, fixed, srv, hit, kind, ttl, val, l0, l1);
    return true;
}

// ---- resolve (lib/server.js:136-429) -------------------------------------------------------
__device__ void resolve_forward(const Params& P, Res& r, uint32_t qidx, uint32_t fixed) {
    const uint8_t* nm = r.p + 12;
    const EngineConst* E = P.eng;
    const bool srv = r.qtype == QT_SRV;
    uint32_t d_off = 0, d_end = r.qn_len - 1;
    uint32_t l0 = 0, l1 = 0;
    r.trunc = 0;
    if (srv) {
        // /^(_[^_.]*)[.](_[^_.]*)[.](.*)/ on query.name() (:141-154); labels hold no '.' here
        l0 = nm[0];
        if (l0 == 0 || nm[1] != '_') { r.rcode = RC_REFUSED; return; }
        for (uint32_t i = 2; i <= l0; i++) if (nm[i] == '_') { r.rcode = RC_REFUSED; return; }
        uint32_t p1 = 1 + l0; l1 = nm[p1];
        if (l1 == 0 || nm[p1 + 1] != '_') { r.rcode = RC_REFUSED; return; }
        for (uint32_t i = 2; i <= l1; i++) if (nm[p1 + i] == '_') { r.rcode = RC_REFUSED; return; }
        d_off = p1 + 1 + l1;
        if (nm[d_off] == 0) { r.rcode = RC_REFUSED; return; }                 // no third part
        // group 3 stops at the first \n or \r (JS '.' excludes line terminators, no '$')
        uint32_t nlp = d_off;
        for (uint32_t pos = d_off; pos < d_end; pos++) {
            if (pos == nlp) { nlp = pos + 1 + nm[pos]; continue; }
            if (nm[pos] == '\n' || nm[pos] == '\r') { d_end = pos; r.trunc = 1; break; }
        }
        if (d_end - d_off - 1 < 1 || d_end <= d_off + 1) { r.rcode = RC_REFUSED; return; }   // :144
    }
    r.d_off = (uint16_t)d_off; r.d_end = (uint16_t)d_end;
    if (d_end <= d_off + 1 && !srv) {                                         // root name: ''
        // isSuffix('.dom', '') is false -> refused; with no dnsDomain: length < 1 -> refused (:198)
        r.rcode = P.ready || E->suffix_len ? RC_REFUSED : RC_SERVFAIL; return;
    }
    // one pass over the domain in dotted view: suffix gate (:157-166, case-sensitive), charset
    // after toLowerCase (:207-215), and where an owner-name pointer may land
    const uint32_t dl = d_end - d_off - 1;
    const uint32_t sl = E->suffix_len;
    bool suffix_ok = sl == 0 || dl >= sl, charset_ok = true, need_b = false;
    uint32_t pos0 = d_end - sl, ptr_tgt = d_off, lastlen = d_off;
    {
        uint32_t nlp = d_off + 1 + nm[d_off];
        for (uint32_t pos = d_off + 1; pos < d_end; pos++) {
            uint32_t c, raw;
            if (pos == nlp) { raw = c = '.'; nlp = pos + 1 + nm[pos]; lastlen = pos; if (need_b) { ptr_tgt = pos; need_b = false; } }
            else {
                raw = nm[pos]; c = lower8(raw);
                if (raw != c) need_b = true;
                if (!((c - 'a' < 26u) || (c - '0' < 10u) || c == '_' || c == '-')) charset_ok = false;
            }
            if (suffix_ok && sl && pos >= pos0 && raw != E->suffix[pos - pos0]) suffix_ok = false;
        }
    }
    if (!suffix_ok) { r.rcode = RC_REFUSED; return; }
    if (!P.ready && !P.route) { r.rcode = RC_SERVFAIL; return; }             // :186-192
    if (!charset_ok) { r.rcode = RC_REFUSED; return; }
    r.ptr_tgt = (need_b || r.trunc) ? (uint16_t)NONE16 : (uint16_t)ptr_tgt;
    r.lastlen = (uint16_t)lastlen;

    FwdKey kg; kg.nm = nm; kg.d_off = d_off; kg.d_end = d_end;
    uint32_t kind = 0, ttl = 0, val = 0;
    const bool hit = probe(P, r, NS_FORWARD, kg, kind, ttl, val);
    finish_forward(P, r, qidx, fixed, srv, hit, kind, ttl, val, l0, l1);
}

// ---- resolvePtr (lib/server.js:67-134) -----------------------------------------------------
__device__ void resolve_ptr(const Params& P, Res& r, uint32_t fixed) {
    const uint8_t* nm = r.p + 12;
    uint32_t nlab = 0, last = 0, prev = 0;
    for (uint32_t q = 0; nm[q]; q += 1 + nm[q]) { prev = last; last = q; ++nlab; }
    // parts.reverse(): [0] must be 'arpa', [1] 'in-addr' — case-sensitive (:71-78)
    bool ok = nlab >= 2 && nm[last] == 4 && nm[last + 1] == 'a' && nm[last + 2] == 'r' && nm[last + 3] == 'p' && nm[last + 4] == 'a' &&
              nm[prev] == 7 && nm[prev + 1] == 'i' && nm[prev + 2] == 'n' && nm[prev + 3] == '-' && nm[prev + 4] == 'a' &&
              nm[prev + 5] == 'd' && nm[prev + 6] == 'd' && nm[prev + 7] == 'r';
    if (!ok) { r.rcode = RC_REFUSED; return; }
    if (!P.ready && !P.route) { r.rcode = RC_SERVFAIL; return; }              // :86-92
    RevKey kg; kg.nm = nm; kg.nlab = nlab - 2; kg.measure();
    uint32_t kind = 0, ttl = 0, val = 0;
    bool hit = kg.length() > 0 && probe(P, r, NS_REVERSE, kg, kind, ttl, val);
    if (!hit) {                                                               // :107-121
        if (P.recursion && r.rd) {                                            // a PTR miss asks every datacenter (:346-354)
            if (P.recursion == 2 && !P.eng->rf_ptr) { r.rcode = RC_REFUSED; return; }
            r.status = ST_MISS; r.rk = RK_NONE; r.rlen = 0; return;
        }
        r.rcode = RC_REFUSED; return;
    }
    if (kind != K_PTR) { r.rcode = RC_SERVFAIL; return; }                     // contract
    r.ttl = ttl; r.val = val; r.rcode = RC_NOERROR; r.rk = RK_PTR;
    size_single(r, fixed, 2 + 10 + P.arena[val]);
}

// onQuery (lib/server.js:471-507) + sizing.  Leaves r ready for emit_response().
__device__ void resolve_query(const Params& P, Res& r, uint32_t len, uint32_t qidx, uint32_t s_sfx) {
    r.status = ST_ANSWERED; r.rk = RK_NONE; r.rlen = 0; r.tc = 0; r.keep_ans = r.keep_add = 0; r.nk = 0; r.n_walk = 0;
    r.ptr_tgt = (uint16_t)NONE16; r.trunc = 0; r.perm = 0; r.ttl = r.val = 0; r.d_off = r.d_end = r.lastlen = 0;
    if (!(r.sp ? decode_staged(r.sp, len, r) : decode(r.p, len, r))) { r.status = ST_DROPPED; return; }
    r.maxsz = P.tcp ? (uint16_t)65535 : r.edns ? (uint16_t)min(max((uint32_t)r.adv, 512u), 1200u) : (uint16_t)512;
    const uint32_t fixed = 12 + r.qn_len + 4 + (r.edns ? 11 : 0);
    r.rk = RK_HEADER; r.rlen = (uint16_t)fixed;
    const bool handled = r.opcode == 0 && (r.qtype == QT_A || r.qtype == QT_SRV || r.qtype == QT_PTR);
    if (!handled) { r.rcode = RC_NOTIMP; return; }                            // :500-505
    STAMP(3);
    if (r.sp && r.qtype != QT_PTR && r.qn_len <= 64 && fast_forward(P, r, s_sfx, qidx, fixed)) return;
    const uint8_t* nm = r.p + 12;
    for (uint32_t q = 0; nm[q];) {                                            // DESIGN.md "in-label dots"
        uint32_t l = nm[q];
        for (uint32_t i = 1; i <= l; i++) if (nm[q + i] == '.') { r.rcode = RC_REFUSED; return; }
        q += 1 + l;
    }
    if (r.qtype == QT_PTR) resolve_ptr(P, r, fixed);
    else resolve_forward(P, r, qidx, fixed);
}

// ---- mname encode (DESIGN.md "Wire spec: encode") ------------------------------------------
// OPT echoed when the query carried one: root owner, type 41, udp size 1200, ttl 0, rdlen 0
__constant__ uint8_t c_opt_rr[11] = { 0, 0, QT_OPT, 0x04, 0xB0, 0, 0, 0, 0, 0, 0 };
// The response staging buffer is XOR-swizzled: 16-byte chunk index ^ (128-byte row & 7).  With one
// 64-byte response per lane, word w of every lane would otherwise fall into 2 of the 32 banks (a
// 16-way conflict on every store); swizzled, a warp's stores spread over more banks, and the flush
// (consecutive 16-byte chunks) still reads each row as a permutation of itself.
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// byte emitter of the generic path: writes the response straight to its place in global memory
struct Out {
    uint8_t* o;
    __device__ void u8(uint32_t v) { *o++ = (uint8_t)v; }
    __device__ void u16(uint32_t v) { o[0] = (uint8_t)(v >> 8); o[1] = (uint8_t)v; o += 2; }
    __device__ void u32(uint32_t v) { o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v; o += 4; }
    __device__ void copy(const uint8_t* s, uint32_t n) { for (uint32_t i = 0; i < n; i++) o[i] = s[i]; o += n; }
};
// the domain part, lower-cased, as wire labels up to `stop` (no terminator)
__device__ void put_dom_labels(Out& w, const Res& r, uint32_t stop) {
    const uint8_t* nm = r.p + 12;
    uint8_t* start = w.o;
    for (uint32_t pos = r.d_off; pos < stop; pos++) w.u8(lower8(nm[pos]));    // length bytes (<64) are unaffected
    if (r.trunc && stop > r.lastlen) start[r.lastlen - r.d_off] = (uint8_t)(r.d_end - r.lastlen - 1);
}
__device__ void put_dom_owner(Out& w, const Res& r) {
    if (r.ptr_tgt != NONE16) { put_dom_labels(w, r, r.ptr_tgt); w.u16(0xC000 | (12 + r.ptr_tgt)); }
    else { put_dom_labels(w, r, r.d_end); w.u8(0); }
}
__device__ void put_rr_head(Out& w, uint32_t type, uint32_t ttl, uint32_t rdlen) { w.u16(type); w.u16(1); w.u32(ttl); w.u16(rdlen); }

__device__ void emit_response(const Params& P, const Res& r, uint8_t* dst, uint32_t qidx) {
    Out w; w.o = dst;
    const uint8_t* p = r.p;
    uint32_t an = 0, ns = 0, ar = r.edns ? 1 : 0;
    switch (r.rk) {
    case RK_A1: case RK_PTR: an = r.keep_ans; break;
    case RK_SOA: ns = r.keep_ans; break;
    case RK_SVC_A: case RK_SVC_SRV: an = r.keep_ans; ar += r.keep_add; break;
    }
    w.u8(p[0]); w.u8(p[1]);
    w.u8(0x80 | (r.opcode << 3) | 0x04 | (r.tc ? 0x02 : 0) | r.rd);          // QR AA TC RD
    w.u8(r.rcode);                                                            // RA=0 Z=0
    w.u16(1); w.u16(an); w.u16(ns); w.u16(ar);
    w.copy(p + 12, r.qn_len + 4);                                             // question, verbatim
    const uint8_t* opt = c_opt_rr;
    bool opt_done = !r.edns;
    if (r.rk == RK_A1 && r.keep_ans) {                                        // :299,310
        put_dom_owner(w, r); put_rr_head(w, QT_A, r.ttl, 4); w.u32(r.val);
    } else if (r.rk == RK_PTR && r.keep_ans) {                                // :130
        w.u16(0xC00C); uint32_t tl = P.arena[r.val]; put_rr_head(w, QT_PTR, r.ttl, tl); w.copy(P.arena + r.val + 1, tl);
    } else if (r.rk == RK_SOA && r.keep_ans) {                                // :286-287
        const EngineConst* E = P.eng;
        put_dom_owner(w, r); put_rr_head(w, QT_SOA, r.ttl, E->soa_len + 20);
        w.copy(E->soa, E->soa_len); w.u32(0); w.u32(10); w.u32(10); w.u32(10); w.u32(r.ttl);
    } else if (r.rk == RK_SVC_A || r.rk == RK_SVC_SRV) {
        const bool srv = r.rk == RK_SVC_SRV;
        SvcView sv; sv.open(P.arena, r.val);
        uint32_t left = r.keep_ans;
        for (uint32_t t = 0; t < r.n_walk && left; t++) {
            const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
            if (k->flags & KID_ADDR_NULL) continue;
            if (srv) {                                                        // :396-400
                const uint8_t* ports = (const uint8_t*)(k + 1);
                const uint8_t* kw = ports + 2 * k->nports;
                for (uint32_t c = 0; c < k->nports && left; c++, left--) {
                    w.u16(0xC00C); put_rr_head(w, QT_SRV, r.ttl, 6 + k->wire_len + dom_wire_len(r));
                    w.u16(0); w.u16(10); w.u16(ld16a(ports + 2 * c));
                    w.copy(kw, k->wire_len); put_dom_labels(w, r, r.d_end); w.u8(0);
                }
            } else {                                                          // :411-414
                uint32_t rttl = (k->flags & KID_HAS_RTTL) ? k->rttl : r.ttl;
                if (r.ttl < rttl) rttl = r.ttl;
                put_dom_owner(w, r); put_rr_head(w, QT_A, rttl, 4); w.u32(k->addr); --left;
            }
        }
        if (srv) {
            if (!opt_done) { w.copy(opt, 11); opt_done = true; }
            left = r.keep_add;
            for (uint32_t t = 0; t < r.n_walk && left; t++) {                 // :401-402
                const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
                if (k->flags & KID_ADDR_NULL) continue;
                const uint8_t* kw = (const uint8_t*)(k + 1) + 2 * k->nports;
                uint32_t rttl = (k->flags & KID_HAS_RTTL) ? k->rttl : r.ttl;
                w.copy(kw, k->wire_len); put_dom_owner(w, r); put_rr_head(w, QT_A, rttl, 4); w.u32(k->addr); --left;
            }
        }
    }
    if (!opt_done) w.copy(opt, 11);
}

// ---- word-wise response writer -------------------------------------------------------------
// A byte stream into shared memory at an arbitrary byte address, stored as aligned 32-bit words;
// only the bytes shared with the neighbouring responses (first / last partial word) go out as
// single bytes, so two threads never write the same word.
// MODE 0: plain shared buffer, 1: XOR-swizzled shared staging (buffer 1024-byte aligned, so the
// swizzle applies to the address itself), 2: global memory (gbase + offset).
// HEADCHK: any put may be the one that completes the first word.  Without it the stream must open
// with put4_first(), and every later store is a plain aligned word.
template <int MODE, bool HEADCHK = false>
struct WrT {
    uint32_t base;       // shared address of the buffer (mode 0)
    uint8_t* gbase;      // global destination (mode 2)
    uint32_t wp;         // position of the aligned word being filled: offset (modes 0, 2) or shared address (mode 1)
    uint32_t acc, fill;  // its bytes so far (fill = 0..3 of them)
    uint32_t head;       // bytes of the FIRST word that belong to the previous response (0..3)
    __device__ void begin(uint32_t buf, uint32_t off) {
        gbase = nullptr; head = off & 3u; acc = 0; fill = head;
        if (MODE == 1) { base = 0; wp = buf + off - head; } else { base = buf; wp = off - head; }
    }
    // global: `g` must be 4-byte aligned (the output buffer is 16-byte aligned), off = byte offset in it
    __device__ void begin_global(uint8_t* g, uint32_t off) { base = 0; gbase = g; head = off & 3u; wp = off - head; acc = 0; fill = head; }
    __device__ __forceinline__ void st32(uint32_t pos, uint32_t v) {
        if (MODE == 2) *(uint32_t*)(gbase + pos) = v;
        else if (MODE == 1) sts32(pos ^ ((pos >> 3) & 0x70u), v);
        else sts32(base + pos, v);
    }
    __device__ __forceinline__ void st8(uint32_t pos, uint32_t v) {
        if (MODE == 2) gbase[pos] = (uint8_t)v;
        else if (MODE == 1) sts8(pos ^ ((pos >> 3) & 0x70u), v & 0xFF);
        else sts8(base + pos, v & 0xFF);
    }
    __device__ __forceinline__ void store(uint32_t v) {
        if (HEADCHK && head) { for (uint32_t b = head; b < 4; b++) st8(wp + b, (v >> (8 * b)) & 0xFF); head = 0; }
        else st32(wp, v);
        wp += 4;
    }
    // the first four bytes of the stream: the only word that may be shared with the previous response
    __device__ __forceinline__ void put4_first(uint32_t v) {
        const uint32_t s8 = 8 * head, w = v << s8;
        if (head == 0) st32(wp, w);
        else for (uint32_t b = head; b < 4; b++) st8(wp + b, (w >> (8 * b)) & 0xFF);
        acc = __funnelshift_l(v, 0u, s8);
        wp += 4; head = 0;
    }
    // four bytes in memory order: the word being filled completes, `fill` bytes carry over
    __device__ __forceinline__ void put4(uint32_t v) {
        const uint32_t s8 = 8 * fill;
        store(acc | (v << s8));
        acc = __funnelshift_l(v, 0u, s8);                // v >> (32 - s8), 0 when s8 == 0
    }
    // v holds n (1..4) bytes in memory order (little-endian integer), upper bytes zero
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        const uint32_t s8 = 8 * fill;
        const uint32_t w = acc | (v << s8);
        if (fill + n >= 4) { store(w); acc = __funnelshift_l(v, 0u, s8); fill = fill + n - 4; }
        else { acc = w; fill += n; }
    }
    __device__ void end() { for (uint32_t b = head; b < fill; b++) st8(wp + b, (acc >> (8 * b)) & 0xFF); }
    // n bytes from shared memory (consecutive unaligned words share their aligned halves)
    __device__ void copy(uint32_t src, uint32_t n) {
        const uint32_t b = src & ~3u, sh = (src & 3u) * 8;
        uint32_t prev = lds32(b), i = 0, k = 1;
        for (; i + 4 <= n; i += 4, k++) { const uint32_t nx = lds32(b + 4 * k); put4(__funnelshift_r(prev, nx, sh)); prev = nx; }
        if (i < n) { const uint32_t nx = lds32(b + 4 * k); put(__funnelshift_r(prev, nx, sh) & ((1u << (8 * (n - i))) - 1), n - i); }
    }
};
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// emit_response() with the word-wise writer, for every response shape of a packet staged in shared
// memory (everything except the SRV line-terminator quirk, which keeps the byte emitter).
__device__ __forceinline__ uint32_t bswap16(uint32_t v) { return ((v & 0xFF) << 8) | ((v >> 8) & 0xFF); }

// domain labels [d_off, stop) of the QNAME, lower-cased (length bytes < 64 are unaffected)
template <class W>
__device__ void put_dom_labels_w(W& w, const Res& r, uint32_t stop) {
    const uint32_t n = stop - r.d_off;
    const uint32_t src = r.sp + 12 + r.d_off, b = src & ~3u, sh = (src & 3u) * 8;
    uint32_t prev = lds32(b), i = 0, k = 1;
    for (; i < n; i += 4, k++) {
        const uint32_t nx = lds32(b + 4 * k);
        uint32_t x = __funnelshift_r(prev, nx, sh);
        prev = nx;
        x |= upper_bytes(x) >> 2;
        const uint32_t nb = n - i;
        if (nb >= 4) w.put4(x); else w.put(x & ((1u << (8 * nb)) - 1), nb);
    }
}
template <class W>
__device__ void put_dom_owner_w(W& w, const Res& r) {
    if (r.ptr_tgt != NONE16) {
        if (r.ptr_tgt != r.d_off) put_dom_labels_w(w, r, r.ptr_tgt);
        const uint32_t ptr = 0xC000u | (12u + r.ptr_tgt);
        w.put(bswap16(ptr), 2);
    } else { put_dom_labels_w(w, r, r.d_end); w.put(0, 1); }
}
template <class W>
__device__ __forceinline__ void put_global_bytes(W& w, const uint8_t* s, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) w.put(__ldg(s + i), 1);
}

template <class W>
__device__ void emit_fast(const Params& P, const Res& r, W& w, uint32_t qidx) {
    const uint32_t p = r.sp;
    uint32_t an = 0, ns = 0, ar = r.edns ? 1 : 0;
    switch (r.rk) {
    case RK_A1: case RK_PTR: an = r.keep_ans; break;
    case RK_SOA: ns = r.keep_ans; break;
    case RK_SVC_A: case RK_SVC_SRV: an = r.keep_ans; ar += r.keep_add; break;
    }
    const uint32_t flags = 0x80u | ((uint32_t)r.opcode << 3) | 0x04u | (r.tc ? 0x02u : 0u) | r.rd;
    w.put4_first((ldsu32(p) & 0xFFFFu) | (flags << 16) | ((uint32_t)r.rcode << 24));   // id, QR AA TC RD, rcode
    w.put4(0x00000100u | (bswap16(an) << 16));                                  // QDCOUNT=1, ANCOUNT
    w.put4(bswap16(ns) | (bswap16(ar) << 16));                                  // NSCOUNT, ARCOUNT
    w.copy(p + 12, r.qn_len + 4);                                               // question, verbatim
    bool opt_done = !r.edns;
    if (r.rk == RK_A1 && r.keep_ans) {                                          // :299,310
        const uint32_t bt = bswap32(r.ttl);
        if (r.ptr_tgt == r.d_off) {              // owner is a bare pointer: the 16-byte RR as four whole words
            w.put4(bswap16(0xC000u | (12u + r.ptr_tgt)) | 0x01000000u);        // ptr | TYPE A ...
            w.put4(0x00000100u | (bt << 16));                                    // ... CLASS IN | ttl (high half)
            w.put4((bt >> 16) | 0x04000000u);                                    // ttl (low half) | RDLENGTH 4
            w.put4(bswap32(r.val));
        } else {
            put_dom_owner_w(w, r);
            w.put4(0x01000100u); w.put4(bt); w.put(0x0400u, 2); w.put4(bswap32(r.val));
        }
    } else if (r.rk == RK_PTR && r.keep_ans) {                                  // :130
        const uint32_t tl = P.arena[r.val];
        w.put(0x0CC0u, 2); w.put4(0x01000C00u); w.put4(bswap32(r.ttl)); w.put(bswap16(tl), 2);
        put_global_bytes(w, P.arena + r.val + 1, tl);
    } else if (r.rk == RK_SOA && r.keep_ans) {                                  // :286-287
        const EngineConst* E = P.eng;
        put_dom_owner_w(w, r);
        w.put4(0x01000600u); w.put4(bswap32(r.ttl)); w.put(bswap16(P.soa_len + 20), 2);
        put_global_bytes(w, E->soa, P.soa_len);
        w.put4(0); w.put4(bswap32(10)); w.put4(bswap32(10)); w.put4(bswap32(10)); w.put4(bswap32(r.ttl));
    } else if (r.rk == RK_SVC_A || r.rk == RK_SVC_SRV) {
        const bool srv = r.rk == RK_SVC_SRV;
        SvcView sv; sv.open(P.arena, r.val);
        const uint32_t dwl = dom_wire_len(r);
        uint32_t left = r.keep_ans;
        for (uint32_t t = 0; t < r.n_walk && left; t++) {
            const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
            const uint32_t fl = k->flags;
            if (fl & KID_ADDR_NULL) continue;
            if (srv) {                                                          // :396-400
                const uint32_t np = k->nports, wl = k->wire_len;
                const uint8_t* ports = (const uint8_t*)(k + 1);
                const uint8_t* kwp = ports + 2 * np;
                for (uint32_t c = 0; c < np && left; c++, left--) {
                    w.put(0x0CC0u, 2); w.put4(0x01002100u); w.put4(bswap32(r.ttl)); w.put(bswap16(6 + wl + dwl), 2);
                    w.put4(0x0A000000u);                                        // priority 0, weight 10
                    w.put(bswap16(ld16a(ports + 2 * c)), 2);
                    put_global_bytes(w, kwp, wl);
                    put_dom_labels_w(w, r, r.d_end); w.put(0, 1);
                }
            } else {                                                            // :411-414
                uint32_t rttl = (fl & KID_HAS_RTTL) ? k->rttl : r.ttl;
                if (r.ttl < rttl) rttl = r.ttl;
                put_dom_owner_w(w, r);
                w.put4(0x01000100u); w.put4(bswap32(rttl)); w.put(0x0400u, 2); w.put4(bswap32(k->addr)); --left;
            }
        }
        if (srv) {
            if (!opt_done) { w.put4(0x04290000u); w.put4(0x000000B0u); w.put(0, 3); opt_done = true; }
            left = r.keep_add;
            for (uint32_t t = 0; t < r.n_walk && left; t++) {                   // :401-402
                const KidRec* k = sv.kid(perm_at(r, t, P.seed, qidx));
                const uint32_t fl = k->flags;
                if (fl & KID_ADDR_NULL) continue;
                const uint8_t* kwp = (const uint8_t*)(k + 1) + 2 * k->nports;
                const uint32_t rttl = (fl & KID_HAS_RTTL) ? k->rttl : r.ttl;
                put_global_bytes(w, kwp, k->wire_len);
                put_dom_owner_w(w, r);
                w.put4(0x01000100u); w.put4(bswap32(rttl)); w.put(0x0400u, 2); w.put4(bswap32(k->addr)); --left;
            }
        }
    }
    if (!opt_done) { w.put4(0x04290000u); w.put4(0x000000B0u); w.put(0, 3); }   // OPT: 00 | 00 29 | 04 B0 | ttl 0 | rdlen 0
    w.end();
}

// ---- the kernel ------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef BB_MIN_BLOCKS
#define BB_MIN_BLOCKS 8      /* 64 registers, no spills: 8 tiles (1024 threads) resident per SM */
#endif
// ORDERED: responses packed in query order (tile bases from a decoupled look-back; a tile waits
// for its predecessors' sizes).  !ORDERED ("arrival" packing): a tile claims its output range
// with one atomicAdd and never waits; response i is still out[out_off[i] .. +out_len[i]).
// MULTI: one launch over several receive regions (routed batches, grid.y = source rank): every
// per-batch pointer advances by its stride per region, sizes and shuffle indices come from device memory.
template <bool ORDERED, bool MULTI>
__global__ void __launch_bounds__(T, BB_MIN_BLOCKS) resolve_kernel(const Params P) {
    __shared__ __align__(16) uint8_t s_in[S_IN + 32];
    __shared__ __align__(1024) uint8_t s_out[S_OUT];         // XOR-swizzled (swz()); 1024-aligned: WrT<1> swizzles addresses
    __shared__ uint32_t s_off[T + 1];
    __shared__ uint32_t s_wsum[8];
    __shared__ unsigned long long s_prefix;
    __shared__ __align__(16) uint8_t s_sfx[256];            // dnsDomain as wire labels, right-aligned (EngineConst::wire_tail)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t ry = MULTI ? blockIdx.y : 0;
    const uint8_t* const r_pkts = MULTI ? P.pkts + ry * P.in_stride : P.pkts;
    const uint32_t* const r_pkt_off = MULTI ? (const uint32_t*)((const uint8_t*)P.pkt_off + ry * P.in_stride) : P.pkt_off;
    const uint32_t* const r_n_dev = (MULTI && P.n_dev) ? (const uint32_t*)((const uint8_t*)P.n_dev + ry * P.in_stride) : nullptr;
    const uint32_t* const r_qidx_map = (MULTI && P.qidx_map) ? (const uint32_t*)((const uint8_t*)P.qidx_map + ry * P.in_stride) : nullptr;
    uint8_t* const r_out = MULTI ? P.out + ry * P.out_stride : P.out;
    uint32_t* const r_out_off = MULTI ? (uint32_t*)((uint8_t*)P.out_off + ry * P.off_stride) : P.out_off;
    uint16_t* const r_out_len = MULTI ? (uint16_t*)((uint8_t*)P.out_len + ry * P.len_stride) : P.out_len;
    uint8_t* const r_status = MULTI ? P.status + ry * P.status_stride : P.status;
    uint32_t* const r_miss_idx = MULTI ? (uint32_t*)((uint8_t*)P.miss_idx + ry * P.miss_stride) : P.miss_idx;
    uint32_t* const r_totals = MULTI ? (uint32_t*)((uint8_t*)P.totals + ry * P.totals_stride) : P.totals;
    unsigned long long* const r_desc = MULTI ? (unsigned long long*)((uint8_t*)P.desc + ry * P.desc_stride) : P.desc;
    uint32_t* const r_counter = MULTI ? (uint32_t*)((uint8_t*)P.counter + ry * P.desc_stride) : P.counter;
    uint32_t* const r_qidx_out = (MULTI && P.qidx_out) ? (uint32_t*)((uint8_t*)P.qidx_out + ry * P.off_stride) : nullptr;
    uint8_t* const r_bounce = (MULTI && P.bounce) ? P.bounce + ry * P.out_stride : P.bounce;

    // Tiles are taken in blockIdx order: like CUB's single-pass scan, the look-back below relies on
    // thread blocks being dispatched in increasing blockIdx order (a block only ever waits for
    // lower-numbered blocks, which are resident or finished).
    for (int i = tid; i < 64; i += T) ((uint32_t*)s_sfx)[i] = __ldg((const uint32_t*)P.eng->wire_tail + i);
    const uint32_t tile = blockIdx.x;
    STAMP(0);
    // routed batches: size known only on the device (header: count, bytes, epoch, sender overflow flag)
    const uint32_t n = r_n_dev ? (r_n_dev[3] ? 0u : r_n_dev[0]) : P.n;
    const uint32_t ntiles = (n + T - 1) / T;
    if (tile < ntiles) {
    const uint32_t q0 = tile * T;
    const uint32_t nq = min((uint32_t)T, n - q0);

    // ---- stage this tile's packets ---------------------------------------------------------
    for (int i = tid; i <= (int)nq; i += T) s_off[i] = r_pkt_off[q0 + i];
    __syncthreads();
    STAMP(1);
    const uint32_t b0 = s_off[0], b1 = s_off[nq];
    const uint32_t a0 = b0 & ~15u;
    const bool staged = b1 >= b0 && b1 - a0 <= S_IN;
    if (staged) {
        const uint4* src = (const uint4*)(r_pkts + a0);
        uint4* dst = (uint4*)s_in;
        const uint32_t nv = (b1 - a0 + 15) >> 4;
        for (uint32_t i = tid; i < nv; i += T) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    STAMP(2);
    // ---- parse + lookup + size ----------------------------------------------------------------
    Res r;
    r.status = ST_DROPPED; r.rlen = 0; r.rk = RK_NONE;
    const uint32_t qidx = (r_qidx_map && tid < (int)nq) ? r_qidx_map[q0 + tid] : P.qidx_base + q0 + tid;
    if (tid < (int)nq) {
        const uint32_t o0 = s_off[tid], o1 = s_off[tid + 1];
        if (o1 >= o0 && o1 - o0 <= 65535u) {
            r.p = staged ? s_in + (o0 - a0) : r_pkts + o0;
            r.sp = staged ? (uint32_t)__cvta_generic_to_shared(s_in) + (o0 - a0) : 0u;
            resolve_query(P, r, o1 - o0, qidx, (uint32_t)__cvta_generic_to_shared(s_sfx));
        }
    }
    const uint32_t my_len = r.rlen;
    const uint32_t my_miss = (tid < (int)nq && r.status == ST_MISS) ? 1u : 0u;

    STAMP(6);
    // ---- CTA scan of (bytes, misses) ------------------------------------------------------------
    uint32_t v = my_len | (my_miss << 24);     // 128 x 1232 < 2^24
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, tot = 0;
    for (int w = 0; w < T / 32; w++) { uint32_t x = s_wsum[w]; if (w < warp) wbase += x; tot += x; }
    const uint32_t excl = wbase + inc - v;
    const uint32_t my_o = excl & 0xFFFFFF, my_mrank = excl >> 24;
    const uint32_t tile_bytes = tot & 0xFFFFFF, tile_miss = tot >> 24;

    STAMP(7);
    // ---- where this tile's responses (and misses) go ------------------------------------------
    if (!ORDERED) {
        if (tid == 0)                                    // one claim per tile: bytes | misses << 40
            s_prefix = atomicAdd(r_desc + P.ntiles_cap, (unsigned long long)tile_bytes | ((unsigned long long)tile_miss << D_MISS_SHIFT));
    } else
    if (warp == 0) {
        volatile unsigned long long* D = r_desc;
        const uint64_t agg = (uint64_t)tile_bytes | ((uint64_t)tile_miss << D_MISS_SHIFT);
        uint64_t ex = 0;
        if (tile == 0) { if (lane == 0) D[0] = D_FLAG_P | agg; }
        else {
            if (lane == 0) D[tile] = D_FLAG_A | agg;
            int base = (int)tile - 1;
            for (;;) {                                   // window of 128 predecessors, 4 independent loads per lane
                uint64_t d[4];
                bool ready;
                do {
                    ready = true;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int idx = base - lane - 32 * k;
                        d[k] = idx >= 0 ? D[idx] : D_FLAG_P;
                        ready &= (d[k] >> 62) != 0;
                    }
                } while (__any_sync(0xffffffffu, !ready));
                uint32_t mypos = 0xFFFFFFFFu;            // distance of the nearest predecessor with an inclusive prefix
#pragma unroll
                for (int k = 3; k >= 0; k--) if ((d[k] >> 62) == 2) mypos = (uint32_t)(lane + 32 * k);
                uint32_t minpos = mypos;
                for (int o = 16; o; o >>= 1) minpos = min(minpos, __shfl_xor_sync(0xffffffffu, minpos, o));
                uint64_t part = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) if ((uint32_t)(lane + 32 * k) <= minpos) part += d[k] & D_VAL;
                ex += warp_sum64(part);
                if (minpos != 0xFFFFFFFFu) break;
                base -= 128;
            }
            if (lane == 0) D[tile] = D_FLAG_P | (ex + agg);
        }
        if (lane == 0) s_prefix = ex;
    }
    __syncthreads();
    STAMP(8);
    const uint64_t ex = s_prefix;
    const uint64_t gbase = ex & ((1ull << D_MISS_SHIFT) - 1);
    const uint32_t mbase = (uint32_t)(ex >> D_MISS_SHIFT);
    const bool overflow = gbase + tile_bytes > (uint64_t)P.out_cap;

    // ---- per-query outputs ---------------------------------------------------------------------
    if (tid < (int)nq) {
        r_out_off[q0 + tid] = (uint32_t)(gbase + my_o);
        r_out_len[q0 + tid] = (uint16_t)my_len;
        r_status[q0 + tid] = r.status;
        if (my_miss) r_miss_idx[mbase + my_mrank] = q0 + tid;
        if (MULTI && r_qidx_out) r_qidx_out[q0 + tid] = qidx;
    }
    if (overflow && tid == 0) r_totals[2] = P.epoch;

    // ---- emit ---------------------------------------------------------------------------------------
    // A tile whose responses fit the staging window is assembled in (swizzled) shared memory and
    // flushed with aligned 16-byte stores.  A larger tile (service answers are ~250 B each), or one
    // with a query on the generic byte path, writes straight to global memory instead: every thread's
    // response is one long contiguous run, and L2 merges its 4-byte stores.
    const bool generic_emit = my_len && !(r.sp && !r.trunc);
    const bool direct = __syncthreads_or(generic_emit) || tile_bytes > (uint32_t)CAPW;
    if (!overflow && direct) {
        // `out` may be pinned host memory (zero-copy results): 4-byte stores over PCIe would be ruinous,
        // so such tiles assemble in the device bounce buffer and then move their contiguous range with
        // coalesced 16-byte stores (the bytes are still in L2)
        uint8_t* const dst = r_bounce ? r_bounce : r_out;
        if (my_len) {
            if (generic_emit) emit_response(P, r, dst + gbase + my_o, qidx);
            else { WrT<2> w; w.begin_global(dst, (uint32_t)(gbase + my_o)); emit_fast(P, r, w, qidx); }
        }
        STAMP(9);
        if (r_bounce && tile_bytes) {
            __syncthreads();
            const uint8_t* src = r_bounce + gbase;
            uint8_t* g = r_out + gbase;
            uint32_t head = (uint32_t)((16 - (gbase & 15)) & 15);
            if (head > tile_bytes) head = tile_bytes;
            if (tid < (int)head) g[tid] = src[tid];
            const uint32_t nv = (tile_bytes - head) >> 4;
            for (uint32_t i = tid; i < nv; i += T) *(uint4*)(g + head + 16 * i) = *(const uint4*)(src + head + 16 * i);
            const uint32_t x0 = head + (nv << 4);
            if (x0 + tid < tile_bytes) g[x0 + tid] = src[x0 + tid];
        }
    } else if (!overflow && tile_bytes) {
        const uint32_t shift = (uint32_t)(gbase & 15);                       // same 16-byte phase in shared and global memory
        if (my_len) { WrT<1> w; w.begin((uint32_t)__cvta_generic_to_shared(s_out), shift + my_o); emit_fast(P, r, w, qidx); }
        __syncthreads();
        STAMP(9);
        uint8_t* g = r_out + gbase;                                           // g[x] <-> s_out[swz(shift + x)]
        uint32_t x0 = 0;
        const uint32_t x1 = tile_bytes;
        uint32_t head = (uint32_t)((16 - (gbase & 15)) & 15);                 // up to 16-byte alignment of the global address
        if (head > x1) head = x1;
        if (tid < (int)head) g[tid] = s_out[swz(shift + tid)];
        x0 = head;
        const uint32_t nv = (x1 - x0) >> 4;
        for (uint32_t i = tid; i < nv; i += T) *(uint4*)(g + x0 + 16 * i) = *(const uint4*)(s_out + swz(shift + x0 + 16 * i));
        x0 += nv << 4;
        if (x0 + tid < x1) g[x0 + tid] = s_out[swz(shift + x0 + tid)];
    }

    STAMP(10);
    }   // tile < ntiles
    // ---- self-cleaning: the last block to finish publishes the totals and resets the placement
    // state for the next launch (blocks beyond ntiles only take part in this count) ------------
    if (warp == 0) {
        uint32_t last = 0;
        // ORDERED: this tile's descriptor store must be visible before it counts itself done.  Arrival
        // packing needs no fence here: the claim was an atomic whose return value this thread already
        // consumed, so it has been performed at L2.
        if (lane == 0) { if (ORDERED) __threadfence(); last = atomicAdd(r_counter + 1, 1u) == gridDim.x - 1; }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {                                // every tile has finished reading descriptors / claiming
            __threadfence();
            volatile unsigned long long* D = r_desc;
            unsigned long long cur = 0;
            if (ORDERED) { if (ntiles) cur = D[ntiles - 1] & D_VAL; }            // inclusive prefix of the last tile
            else cur = D[P.ntiles_cap];
            if (lane == 0) {
                const uint32_t tb = (uint32_t)(cur & ((1ull << D_MISS_SHIFT) - 1));
                r_out_off[n] = tb; r_totals[0] = tb; r_totals[1] = (uint32_t)(cur >> D_MISS_SHIFT); r_totals[3] = P.epoch;
                if (MULTI) {                   // what a host reading only the totals needs to know about the region
                    r_totals[4] = n; r_totals[5] = r_n_dev ? r_n_dev[3] : 0u; r_totals[6] = P.err_in ? *(volatile const uint32_t*)P.err_in : 0u;
                }
            }
            __syncwarp();
            if (ORDERED) { for (uint32_t i = lane; i < ntiles; i += 32) r_desc[i] = 0; }
            else if (lane == 0) r_desc[P.ntiles_cap] = 0;
            if (lane == 0) r_counter[1] = 0;
        }
    }
}


// ---- multi-GPU: route + push ----------------------------------------------------------------
// Ingress side of the sharded design.  Each query is parsed just far enough to
// know its lookup key (the same code as resolve, in route mode) and is then written straight
// into the receive region (this rank -> owner rank) in the OWNER's HBM with peer-to-peer stores
// over NVLink: no staging buffer, no separate collective.  Space inside a region is claimed
// with sender-local atomics (a region has exactly one writer rank), so nothing atomic ever
// crosses the link.  A region is laid out as an ordinary batch (packed packets + u32 offsets)
// plus the original query index of each packet, so the owner resolves it with resolve_kernel.
constexpr int MAX_RANKS = 8;
constexpr int PUSH_CNT_SHIFT = 40;
struct PushParams {
    Params P;                              // pkts, pkt_off, n, eng, route = 1, nranks, rank
    uint8_t* region[MAX_RANKS];            // region (rank -> d) inside rank d's receive buffer (peer-mapped)
    uint32_t cap_q, cap_b;                 // capacity of one region: queries, packet bytes
    unsigned long long* cursor;            // [nranks], local: count << 40 | bytes
    uint32_t* done;                        // finished-block counter
    uint32_t* err;                         // set when a region overflows
    uint32_t qidx_base, epoch;
};
__host__ __device__ inline size_t region_off_array(uint32_t) { return 16; }
__host__ __device__ inline size_t region_qidx_array(uint32_t cap_q) { return 16 + 4 * ((size_t)cap_q + 1); }
__host__ __device__ inline size_t region_bytes(uint32_t cap_q) { return (16 + 4 * ((size_t)cap_q + 1) + 4 * (size_t)cap_q + 15) & ~(size_t)15; }
__host__ __device__ inline size_t region_size(uint32_t cap_q, uint32_t cap_b) { return (region_bytes(cap_q) + cap_b + 64 + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(T, BB_MIN_BLOCKS) route_push_kernel(const PushParams A) {
    const Params& P = A.P;
    __shared__ __align__(16) uint8_t s_in[S_IN + 32];
    __shared__ __align__(16) uint8_t s_sorted[S_IN + 16 * MAX_RANKS + 32];     // the tile's packets grouped by owner
    __shared__ uint32_t s_moff[T], s_mq[T];                                    // per-owner-grouped offsets / query indices
    __shared__ uint32_t s_off[T + 1];
    // per (tile, owner): queries << 24 | packet bytes.  One 32-bit word so that the claim is a native shared-memory
    // add (a 64-bit one is a compare-and-swap loop, and the whole tile contends on nranks words).
    // <= 128 queries of <= 65535 bytes each: the byte field stays below 2^24.
    __shared__ uint32_t s_cur[MAX_RANKS];
    __shared__ unsigned long long s_base[MAX_RANKS];
    __shared__ uint32_t s_kstart[MAX_RANKS + 1], s_bstart[MAX_RANKS];
    __shared__ uint32_t s_ovf;
    __shared__ __align__(16) uint8_t s_sfx[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 64; i += T) ((uint32_t*)s_sfx)[i] = __ldg((const uint32_t*)P.eng->wire_tail + i);
    if (tid < MAX_RANKS) s_cur[tid] = 0;
    if (tid == 0) s_ovf = 0;
    STAMP(0);
    const uint32_t q0 = blockIdx.x * T;
    const uint32_t nq = min((uint32_t)T, P.n - q0);
    for (int i = tid; i <= (int)nq; i += T) s_off[i] = P.pkt_off[q0 + i];
    __syncthreads();
    STAMP(1);
    const uint32_t b0 = s_off[0], b1 = s_off[nq];
    const uint32_t a0 = b0 & ~15u;
    const bool staged = b1 >= b0 && b1 - a0 <= S_IN;
    if (staged) {
        const uint4* src = (const uint4*)(P.pkts + a0);
        uint4* dst = (uint4*)s_in;
        const uint32_t nv = (b1 - a0 + 15) >> 4;
        for (uint32_t i = tid; i < nv; i += T) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    STAMP(2);
    Res r;
    r.owner = (uint8_t)P.rank;                     // queries that need no lookup are answered where they arrived
    r.sp = 0; r.p = nullptr;
    uint32_t len = 0, k = 0, boff = 0;
    const bool have = tid < (int)nq;
    if (have) {
        const uint32_t o0 = s_off[tid], o1 = s_off[tid + 1];
        if (o1 >= o0 && o1 - o0 <= 65535u) {
            len = o1 - o0;
            r.p = staged ? s_in + (o0 - a0) : P.pkts + o0;
            r.sp = staged ? (uint32_t)__cvta_generic_to_shared(s_in) + (o0 - a0) : 0u;
            resolve_query(P, r, len, 0, (uint32_t)__cvta_generic_to_shared(s_sfx));
            if (r.owner >= P.nranks) r.owner = (uint8_t)P.rank;
        }
        const uint32_t old = atomicAdd(&s_cur[r.owner], (1u << 24) | len);
        k = old >> 24; boff = old & 0xFFFFFFu;
    }
    __syncthreads();
    STAMP(6);
    // one claim per (tile, owner) in the sender-local cursor of region (this rank -> owner)
    if (tid < (int)P.nranks) {
        const uint32_t t = s_cur[tid];
        s_base[tid] = t ? atomicAdd(A.cursor + tid, ((unsigned long long)(t >> 24) << PUSH_CNT_SHIFT) | (t & 0xFFFFFFu)) : 0ull;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t kk = 0, bb = 0;
        for (uint32_t d = 0; d < P.nranks; d++) {
            const uint32_t t = s_cur[d]; const unsigned long long base = s_base[d];
            const uint32_t cnt = t >> 24, nb = t & 0xFFFFFFu;
            const uint32_t gk = (uint32_t)(base >> PUSH_CNT_SHIFT);
            const unsigned long long gb = base & ((1ull << PUSH_CNT_SHIFT) - 1);
            if (cnt && (gk + cnt > A.cap_q || gb + nb > A.cap_b)) { s_ovf = 1; *A.err = 1; }
            s_kstart[d] = kk; kk += cnt;
            bb += (uint32_t)((gb - bb) & 15);      // group start has the 16-byte phase of its destination
            s_bstart[d] = bb; bb += nb;
        }
        s_kstart[P.nranks] = kk;
    }
    __syncthreads();
    const bool ovf = s_ovf != 0;
    STAMP(7);
    if (have && !ovf) {
        const unsigned long long base = s_base[r.owner];
        const uint32_t gb = (uint32_t)(base & ((1ull << PUSH_CNT_SHIFT) - 1)) + boff;
        if (staged) {
            s_moff[s_kstart[r.owner] + k] = gb;
            s_mq[s_kstart[r.owner] + k] = A.qidx_base + q0 + tid;
            WrT<0, true> w; w.begin((uint32_t)__cvta_generic_to_shared(s_sorted), s_bstart[r.owner] + boff);
            w.copy(r.sp, len);
            w.end();
        } else {                                   // oversized tile: plain peer stores from global memory
            const uint32_t gk = (uint32_t)(base >> PUSH_CNT_SHIFT) + k;
            uint8_t* reg = A.region[r.owner];
            ((uint32_t*)(reg + region_off_array(A.cap_q)))[gk] = gb;
            ((uint32_t*)(reg + region_qidx_array(A.cap_q)))[gk] = A.qidx_base + q0 + tid;
            uint8_t* dst = reg + region_bytes(A.cap_q) + gb;
            const uint8_t* src = P.pkts + s_off[tid];
            for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
        }
    }
    __syncthreads();
    STAMP(8);
    if (staged && !ovf) {
        // per owner: one contiguous chunk of packets and of metadata, pushed with coalesced peer stores
        for (uint32_t d = 0; d < P.nranks; d++) {
            const uint32_t t = s_cur[d]; const unsigned long long base = s_base[d];
            const uint32_t cnt = t >> 24, nb = t & 0xFFFFFFu;
            if (!cnt) continue;
            uint8_t* reg = A.region[d];
            const uint32_t gk = (uint32_t)(base >> PUSH_CNT_SHIFT);
            const unsigned long long gb = base & ((1ull << PUSH_CNT_SHIFT) - 1);
            uint32_t* go = (uint32_t*)(reg + region_off_array(A.cap_q)) + gk;
            uint32_t* gq = (uint32_t*)(reg + region_qidx_array(A.cap_q)) + gk;
            for (uint32_t i = tid; i < cnt; i += T) { go[i] = s_moff[s_kstart[d] + i]; gq[i] = s_mq[s_kstart[d] + i]; }
            uint8_t* g = reg + region_bytes(A.cap_q) + gb;                    // g[x] <-> s_sorted[s_bstart[d] + x]
            const uint8_t* sm = s_sorted + s_bstart[d];
            uint32_t x0 = 0;
            uint32_t head = (uint32_t)((16 - (gb & 15)) & 15);
            if (head > nb) head = nb;
            if (tid < (int)head) g[tid] = sm[tid];
            x0 = head;
            const uint32_t nv = (nb - x0) >> 4;
            for (uint32_t i = tid; i < nv; i += T) *(uint4*)(g + x0 + 16 * i) = *(const uint4*)(sm + x0 + 16 * i);
            x0 += nv << 4;
            if (x0 + tid < nb) g[x0 + tid] = sm[x0 + tid];
        }
    }
    // The last block publishes the region headers (count, bytes, end-of-offsets sentinel) and then,
    // after a system-scope fence, the epoch flag the owner's wait kernel spins on: the exchange
    // needs no collective, only this ordered pair of peer stores per (source, owner).
    // A block orders its peer stores before its done count with a device-scope fence only (the barrier
    // orders every thread's stores before thread 0's fence; fences are cumulative).  The system-scope
    // fence is the last block's alone: it has observed every other block's count, so by causality
    // order all their stores precede its flag store for whoever acquires the flag at system scope.
    __syncthreads();
    STAMP(9);
    if (warp == 0) {
        uint32_t last = 0;
        if (lane == 0) { __threadfence(); last = atomicAdd(A.done, 1u) == gridDim.x - 1; }
        STAMP(10);
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && lane < (int)P.nranks) {
            __threadfence();
            const unsigned long long c = *(volatile unsigned long long*)(A.cursor + lane);
            const uint32_t cnt = min((uint32_t)(c >> PUSH_CNT_SHIFT), A.cap_q);
            const uint32_t nb = (uint32_t)min(c & ((1ull << PUSH_CNT_SHIFT) - 1), (unsigned long long)A.cap_b);
            uint8_t* reg = A.region[lane];
            ((uint32_t*)(reg + region_off_array(A.cap_q)))[cnt] = nb;
            volatile uint32_t* hdr = (volatile uint32_t*)reg;
            hdr[0] = cnt; hdr[1] = nb; hdr[3] = *A.err;
            __threadfence_system();
            hdr[2] = A.epoch;                      // the flag: everything above is visible to whoever sees it
            if (P.stage_log && lane == 0) P.stage_log[(size_t)gridDim.x * NSTAGE] = gtime();   // one extra row: flag published
            A.cursor[lane] = 0;
        }
        if (last && lane == 0) *A.done = 0;
    }
}

// Owner side: wait (bounded) until every source rank has published `epoch` in its region header.
__global__ void wait_regions_kernel(const uint8_t* recv_set, size_t reg_size, uint32_t nranks, uint32_t epoch, uint32_t* err) {
    const uint32_t r = threadIdx.x;
    if (r < nranks) {
        const volatile uint32_t* hdr = (const volatile uint32_t*)(recv_set + (size_t)r * reg_size);
        unsigned long long spins = 0;
        while (hdr[2] != epoch) {
            if (++spins > (1ull << 28)) { *err = 2; break; }      // ~seconds: a peer died; fail instead of hanging
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

// Incremental zone update: overwrite the listed slots (one thread per 16-byte chunk).  Runs with no
// batch in flight (bb_engine_apply_update synchronises first), so a reader never sees half a slot.
__global__ void patch_slots_kernel(Slot* table, const uint32_t* idx, const uint4* data, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4u * n) ((uint4*)(table + idx[t >> 2]))[t & 3u] = data[t];
}

}  // namespace bbk

// =================================================================================================
// host side of the C ABI
// =================================================================================================
static thread_local std::string g_cuda_err;
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { g_cuda_err = std::string(#call) + ": " + cudaGetErrorString(_e); return BB_ERR_CUDA; } } while (0)

namespace {
constexpr int NSLOTS = 4;
struct SlotCtx {
    cudaStream_t stream = nullptr; cudaEvent_t ev = nullptr;
    uint8_t* d_pkts = nullptr; uint32_t* d_off = nullptr; uint8_t* d_out = nullptr; uint32_t* d_out_off = nullptr;
    uint8_t* d_status = nullptr; uint32_t* d_miss = nullptr; uint32_t* d_totals = nullptr; uint16_t* d_out_len = nullptr;
    unsigned long long* d_desc = nullptr;        // [ntiles_max] + counter
    uint32_t* h_totals = nullptr;                // pinned
    // pending call
    bool busy = false, zero_copy = false; uint32_t n = 0, epoch = 0;
    const void* zc_seen[5] = {}; bool zc_ok = false;        // last output pointer set checked for being pinned
    uint8_t* out = nullptr; uint32_t out_cap = 0; uint32_t* miss_idx = nullptr; uint32_t* n_miss = nullptr;
};
}

struct bb_engine {
    bb::EngineConst hconst; bb::EngineConst* d_const = nullptr;
    bb::Slot* d_table = nullptr; uint8_t* d_arena = nullptr; uint32_t mask = 0; int ready = 0;
    int device = 0, ordered = 0; uint32_t max_batch = 0, max_bytes = 0, out_dev_cap = 0, max_tiles = 0;
    SlotCtx slots[NSLOTS];
    // look-back / claim scratch of bb_resolve_batch_device, one per caller stream (launches on
    // different streams may overlap)
    static constexpr int MAX_DEV_STREAMS = 16;
    void* dev_stream[MAX_DEV_STREAMS] = {}; unsigned long long* dev_desc[MAX_DEV_STREAMS] = {}; int n_dev_streams = 0;
    uint64_t launches = 0, epoch = 0;
    unsigned long long* stage_log = nullptr;
    // the zone this engine last synchronised with (incremental updates only continue from that state)
    const bb_zone* zone = nullptr; uint64_t zone_gen = 0; uint64_t arena_cap = 0;
};

static bool name_to_wire(const std::string& s, std::string& out) {
    out.clear(); size_t st = 0;
    if (s.empty()) return false;
    for (size_t i = 0; i <= s.size(); i++) if (i == s.size() || s[i] == '.') {
        size_t l = i - st; if (l < 1 || l > 63) return false;
        out.push_back((char)l); out.append(s, st, l); st = i + 1;
    }
    return true;
}

extern "C" {

const char* bb_strerror(int err) {
    switch (err) {
    case BB_OK: return "ok";
    case BB_ERR_ARG: return "invalid argument";
    case BB_ERR_SNAPSHOT: return "snapshot is not valid JSON-lines (or repeats a path)";
    case BB_ERR_CUDA: return "CUDA error";
    case BB_ERR_NOMEM: return "out of memory";
    case BB_ERR_CAPACITY: return "output capacity too small for this batch";
    case BB_ERR_NO_DEVICE: return "no CUDA device (binder_b200 has no CPU fallback)";
    case BB_ERR_DOMAIN: return "dns_domain must be a lower-case, encodable DNS name";
    case BB_ERR_PROTOCOL: return "balancer frame stream: unknown frame type, INBOUND_TCP, or oversized packet";
    }
    return "unknown error";
}
const char* bb_last_cuda_error(void) { return g_cuda_err.c_str(); }
int bb_abi_version(void) { return BB_ABI_VERSION; }

void* bb_host_alloc(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
void bb_host_free(void* p) { if (p) cudaFreeHost(p); }

static int engine_alloc(bb_engine* e) {
    CK(cudaSetDevice(e->device));
    CK(cudaMalloc(&e->d_const, sizeof(bb::EngineConst)));
    CK(cudaMemcpy(e->d_const, &e->hconst, sizeof(bb::EngineConst), cudaMemcpyHostToDevice));
    e->max_tiles = (e->max_batch + bbk::T - 1) / bbk::T;
    uint64_t cap = (uint64_t)e->max_batch * 512;                 // device-side response buffer per slot
    if (cap < (1u << 20)) cap = 1u << 20;
    if (cap > 0xFFFFFF00ull) cap = 0xFFFFFF00ull;
    e->out_dev_cap = (uint32_t)cap;
    for (auto& s : e->slots) {
        CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        CK(cudaMalloc(&s.d_pkts, (size_t)e->max_bytes + 64));
        CK(cudaMalloc(&s.d_off, ((size_t)e->max_batch + 1) * 4));
        CK(cudaMalloc(&s.d_out, (size_t)e->out_dev_cap + 64));
        CK(cudaMalloc(&s.d_out_off, ((size_t)e->max_batch + 1) * 4));
        CK(cudaMalloc(&s.d_status, (size_t)e->max_batch + 16));
        CK(cudaMalloc(&s.d_out_len, (size_t)e->max_batch * 2 + 16));
        CK(cudaMalloc(&s.d_miss, (size_t)e->max_batch * 4 + 16));
        CK(cudaMalloc(&s.d_totals, 16));
        CK(cudaMalloc(&s.d_desc, ((size_t)e->max_tiles + 4) * 8));
        CK(cudaMemset(s.d_desc, 0, ((size_t)e->max_tiles + 4) * 8));
        CK(cudaMemset(s.d_totals, 0, 16));
        CK(cudaMallocHost(&s.h_totals, 16));
    }
    return BB_OK;
}

bb_engine* bb_engine_create(const bb_engine_opts* o, int* err) {
    auto fail = [&](int c) -> bb_engine* { if (err) *err = c; return nullptr; };
    if (err) *err = BB_OK;
    if (!o || !o->dns_domain) return fail(BB_ERR_ARG);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(BB_ERR_NO_DEVICE);
    if (o->device < 0 || o->device >= ndev) return fail(BB_ERR_ARG);
    std::string dom = o->dns_domain, w, hw;
    for (char c : dom) if (c >= 'A' && c <= 'Z') return fail(BB_ERR_DOMAIN);
    if (!name_to_wire(dom, w) || !name_to_wire("hostmaster." + dom, hw) || hw.size() + 1 > 255) return fail(BB_ERR_DOMAIN);
    bb_engine* e = new bb_engine();
    memset(&e->hconst, 0, sizeof e->hconst);
    e->hconst.suffix_len = (uint32_t)dom.size() + 1;
    e->hconst.suffix[0] = '.'; memcpy(e->hconst.suffix + 1, dom.data(), dom.size());
    // SOARecord(dnsDomain): mname = dnsDomain, rname = hostmaster.<dnsDomain>, both uncompressed
    memcpy(e->hconst.soa, w.data(), w.size()); e->hconst.soa[w.size()] = 0;
    memcpy(e->hconst.soa + w.size() + 1, hw.data(), hw.size()); e->hconst.soa[w.size() + 1 + hw.size()] = 0;
    e->hconst.soa_len = (uint32_t)(w.size() + 1 + hw.size() + 1);
    memcpy(e->hconst.wire_tail + 256 - w.size(), w.data(), w.size());   // word-wise suffix gate compares the name's tail with this
    e->hconst.recursion = o->recursion ? 1 : 0;
    e->device = o->device; e->ordered = o->ordered_output ? 1 : 0;
    e->max_batch = o->max_batch ? o->max_batch : (1u << 20);
    if (e->max_batch > (1u << 22)) { delete e; return fail(BB_ERR_ARG); }
    e->max_bytes = o->max_batch_bytes ? o->max_batch_bytes : e->max_batch * 64;
    int rc = engine_alloc(e);
    if (rc != BB_OK) { bb_engine_destroy(e); return fail(rc); }
    return e;
}

void bb_engine_destroy(bb_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& s : e->slots) {
        cudaFree(s.d_pkts); cudaFree(s.d_off); cudaFree(s.d_out); cudaFree(s.d_out_off); cudaFree(s.d_status); cudaFree(s.d_out_len);
        cudaFree(s.d_miss); cudaFree(s.d_totals); cudaFree(s.d_desc); if (s.h_totals) cudaFreeHost(s.h_totals);
        if (s.ev) cudaEventDestroy(s.ev); if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (int i = 0; i < e->n_dev_streams; i++) cudaFree(e->dev_desc[i]);
    cudaFree(e->d_const); cudaFree(e->d_table); cudaFree(e->d_arena);
    delete e;
}

int bb_engine_swap_zone(bb_engine* e, const bb_zone* z) {
    if (!e || !z) return BB_ERR_ARG;
    const bb::ZoneImage* img = bb_zone_image(z);
    CK(cudaSetDevice(e->device));
    bb::Slot* nt = nullptr; uint8_t* na = nullptr;
    // head room in the arena: updates append re-derived records (bb_engine_apply_update)
    const uint64_t cap = img->arena_len + (img->arena_len / 2 > (4u << 20) ? img->arena_len / 2 : (4u << 20));
    CK(cudaMalloc(&nt, (size_t)img->nslots * sizeof(bb::Slot)));
    CK(cudaMalloc(&na, (size_t)cap + 64));
    CK(cudaMemcpy(nt, img->slots, (size_t)img->nslots * sizeof(bb::Slot), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(na, img->arena, (size_t)img->arena_len, cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());                 // batches in flight finish on the old epoch
    bb::Slot* ot = e->d_table; uint8_t* oa = e->d_arena;
    e->d_table = nt; e->d_arena = na; e->mask = img->nslots - 1; e->ready = img->ready; e->arena_cap = cap;
    cudaFree(ot); cudaFree(oa);
    // the zone's pending-change list restarts here (a zone feeds one engine incrementally)
    bb_zone_mark_synced(const_cast<bb_zone*>(z));
    e->zone = z; e->zone_gen = bb_zone_sync_gen(z);
    return BB_OK;
}

// After bb_zone_apply: ship what changed — the touched 64-byte slots and the arena tail — instead of the
// whole image.  Falls back to a full swap when the table was laid out again, the arena outgrew its
// device allocation, or this engine is not the one that took the zone's previous changes.
int bb_engine_apply_update(bb_engine* e, bb_zone* z) {
    if (!e || !z) return BB_ERR_ARG;
    const bb::ZoneImage* img = bb_zone_image(z);
    const uint32_t* slots = nullptr; uint32_t n = 0; uint64_t from = 0; int relaid = 0;
    if (bb_zone_pending(z, &slots, &n, &from, &relaid) != BB_OK) return BB_ERR_ARG;
    if (relaid || !e->d_table || e->zone != z || e->zone_gen != bb_zone_sync_gen(z) || e->mask != img->nslots - 1 ||
        img->arena_len > e->arena_cap)
        return bb_engine_swap_zone(e, z);
    CK(cudaSetDevice(e->device));
    CK(cudaDeviceSynchronize());                 // batches in flight finish on the old state
    if (img->arena_len > from)
        CK(cudaMemcpy(e->d_arena + from, img->arena + from, (size_t)(img->arena_len - from), cudaMemcpyHostToDevice));
    if (n) {
        std::vector<bb::Slot> data(n);
        for (uint32_t i = 0; i < n; i++) data[i] = img->slots[slots[i]];
        uint32_t* d_idx = nullptr; uint4* d_data = nullptr;
        CK(cudaMalloc(&d_idx, (size_t)n * 4)); 
        if (cudaMalloc(&d_data, (size_t)n * sizeof(bb::Slot)) != cudaSuccess) { cudaFree(d_idx); g_cuda_err = "cudaMalloc (slot patch)"; return BB_ERR_CUDA; }
        cudaError_t c1 = cudaMemcpy(d_idx, slots, (size_t)n * 4, cudaMemcpyHostToDevice);
        cudaError_t c2 = cudaMemcpy(d_data, data.data(), (size_t)n * sizeof(bb::Slot), cudaMemcpyHostToDevice);
        if (c1 == cudaSuccess && c2 == cudaSuccess) {
            bbk::patch_slots_kernel<<<(4 * n + 255) / 256, 256>>>(e->d_table, d_idx, d_data, n);
            c1 = cudaGetLastError(); c2 = cudaDeviceSynchronize();
        }
        cudaFree(d_idx); cudaFree(d_data);
        if (c1 != cudaSuccess || c2 != cudaSuccess) { g_cuda_err = cudaGetErrorString(c1 != cudaSuccess ? c1 : c2); return BB_ERR_CUDA; }
        e->launches++;
    }
    e->ready = img->ready;
    bb_zone_mark_synced(z);
    e->zone_gen = bb_zone_sync_gen(z);
    return BB_OK;
}
// Recursion pre-filter (lib/recursion.js:329-344).  region_domain =
// Recursion's opts.dnsDomain; dc_names = the keys of self.dcs that still have an upstream after the
// "not one of my own addresses" filter (:360-379); ptr_forwardable = any such upstream exists at all.
// With a filter set, a miss Recursion.resolve() would answer REFUSED without asking anyone is answered
// REFUSED by the kernel and never enters miss_idx.  NULL region_domain removes the filter.
int bb_engine_set_recursion_filter(bb_engine* e, const char* region_domain, const char* const* dc_names, uint32_t n_dc,
                                   int ptr_forwardable) {
    if (!e || (n_dc && !dc_names) || n_dc > bb::RF_MAX_DC) return BB_ERR_ARG;
    bb::EngineConst& C = e->hconst;
    if (!region_domain) {
        if (C.recursion) C.recursion = 1;
    } else {
        if (!C.recursion) return BB_ERR_ARG;                       // the engine was created without recursion
        const size_t L = strlen(region_domain);
        if (L > 255) return BB_ERR_ARG;
        for (uint32_t k = 0; k < n_dc; k++) { const size_t l = dc_names[k] ? strlen(dc_names[k]) : 0; if (l < 1 || l > 63) return BB_ERR_ARG; }
        C.rf_dom_len = (uint32_t)L; memset(C.rf_dom, 0, sizeof C.rf_dom); memcpy(C.rf_dom, region_domain, L);
        C.rf_ndc = n_dc; memset(C.rf_dc, 0, sizeof C.rf_dc); memset(C.rf_dc_len, 0, sizeof C.rf_dc_len);
        for (uint32_t k = 0; k < n_dc; k++) { C.rf_dc_len[k] = (uint8_t)strlen(dc_names[k]); memcpy(C.rf_dc[k], dc_names[k], C.rf_dc_len[k]); }
        C.rf_ptr = ptr_forwardable ? 1 : 0;
        C.recursion = 2;
    }
    CK(cudaSetDevice(e->device));
    CK(cudaDeviceSynchronize());                                   // batches in flight finish with the old filter
    CK(cudaMemcpy(e->d_const, &e->hconst, sizeof(bb::EngineConst), cudaMemcpyHostToDevice));
    return BB_OK;
}
int bb_engine_is_ready(const bb_engine* e) { return e && e->ready; }
int bb_engine_slots(const bb_engine*) { return NSLOTS; }
uint64_t bb_engine_launch_count(const bb_engine* e) { return e ? e->launches : 0; }
uint32_t bb_engine_launch_epoch(const bb_engine* e) { return e ? (uint32_t)e->epoch : 0; }
void bb_engine_set_stage_log(bb_engine* e, unsigned long long* d_log) { if (e) e->stage_log = d_log; }

static int launch(bb_engine* e, unsigned long long* desc, const uint8_t* d_pkts, const uint32_t* d_off, uint32_t n,
                  uint64_t seed, uint32_t qidx_base, uint8_t* d_out, uint32_t out_cap, uint32_t* d_out_off, uint16_t* d_out_len,
                  uint8_t* d_status, uint32_t* d_miss, uint32_t* d_totals, cudaStream_t st, uint8_t* bounce = nullptr, uint32_t flags = 0) {
    bbk::Params P;
    P.pkts = d_pkts; P.pkt_off = d_off; P.n = n; P.seed = seed; P.qidx_base = qidx_base;
    P.out = d_out; P.out_cap = out_cap; P.out_off = d_out_off; P.out_len = d_out_len; P.status = d_status; P.miss_idx = d_miss; P.totals = d_totals;
    P.table = e->d_table; P.mask = e->mask; P.arena = e->d_arena; P.ready = e->ready && e->d_table;
    P.eng = e->d_const; P.suffix_len = e->hconst.suffix_len; P.soa_len = e->hconst.soa_len; P.recursion = e->hconst.recursion;
    P.ntiles = (n + bbk::T - 1) / bbk::T;
    P.desc = desc; P.ntiles_cap = e->max_tiles; P.counter = (uint32_t*)(desc + e->max_tiles + 1);
    P.epoch = (uint32_t)(++e->epoch);
    P.stage_log = e->stage_log;
    P.n_dev = nullptr; P.qidx_map = nullptr; P.route = 0; P.nranks = 1; P.rank = 0; P.regions = 0; P.bounce = bounce; P.qidx_out = nullptr; P.err_in = nullptr; P.tcp = (flags & BB_BATCH_TCP) ? 1u : 0u;
    if (n == 0) { CK(cudaMemsetAsync(d_out_off, 0, 4, st)); CK(cudaMemsetAsync(d_totals, 0, 16, st)); return BB_OK; }
    // no memsets: the kernel leaves desc/counter zeroed for the next launch (self-cleaning)
    if (e->ordered) bbk::resolve_kernel<true, false><<<P.ntiles, bbk::T, 0, st>>>(P);
    else bbk::resolve_kernel<false, false><<<P.ntiles, bbk::T, 0, st>>>(P);
    CK(cudaGetLastError());
    e->launches++;
    return BB_OK;
}

int bb_resolve_batch_device(bb_engine* e, const uint8_t* d_pkts, const uint32_t* d_pkt_off, uint32_t n,
                            uint64_t seed, uint32_t qidx_base, uint8_t* d_out, uint32_t out_cap, uint32_t* d_out_off,
                            uint16_t* d_out_len, uint8_t* d_status, uint32_t* d_miss_idx, uint32_t* d_totals, void* stream) {
    if (!e || n > e->max_batch || ((uintptr_t)d_pkts & 15) || ((uintptr_t)d_out & 15)) return BB_ERR_ARG;
    int si = -1;
    for (int i = 0; i < e->n_dev_streams; i++) if (e->dev_stream[i] == stream) si = i;
    if (si < 0) {
        if (e->n_dev_streams == bb_engine::MAX_DEV_STREAMS) return BB_ERR_ARG;
        CK(cudaSetDevice(e->device));
        unsigned long long* dsc = nullptr;
        CK(cudaMalloc(&dsc, ((size_t)e->max_tiles + 4) * 8));
        CK(cudaMemset(dsc, 0, ((size_t)e->max_tiles + 4) * 8));
        si = e->n_dev_streams++; e->dev_stream[si] = stream; e->dev_desc[si] = dsc;
    }
    return launch(e, e->dev_desc[si], d_pkts, d_pkt_off, n, seed, qidx_base, d_out, out_cap, d_out_off, d_out_len, d_status, d_miss_idx,
                  d_totals, (cudaStream_t)stream);
}

int bb_resolve_submit_ex(bb_engine* e, int slot, const uint8_t* pkts, const uint32_t* pkt_off, uint32_t n, uint64_t seed,
                         uint32_t qidx_base, uint8_t* out, uint32_t out_cap, uint32_t* out_off, uint16_t* out_len, uint8_t* status,
                         uint32_t* miss_idx, uint32_t* n_miss, uint32_t flags) {
    if (!e || slot < 0 || slot >= NSLOTS || !pkt_off || !out_off || !n_miss || (n && (!pkts || !status || !miss_idx || !out_len))) return BB_ERR_ARG;
    SlotCtx& s = e->slots[slot];
    if (s.busy || n > e->max_batch) return BB_ERR_ARG;
    const uint32_t total_in = pkt_off[n];
    if (total_in > e->max_bytes) return BB_ERR_ARG;
    CK(cudaSetDevice(e->device));
    if (total_in) CK(cudaMemcpyAsync(s.d_pkts, pkts, total_in, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.d_off, pkt_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s.stream));
    // Zero-copy results: when every output buffer is pinned (mapped) host memory — bb_host_alloc — the
    // kernel's coalesced flush writes the responses straight into it over PCIe: no device staging, no
    // D2H copies and no host round trip to learn the sizes.  Otherwise: staged copies (below).
    const void* outs[5] = { out, out_off, out_len, status, miss_idx };
    if (memcmp(outs, s.zc_seen, sizeof outs) != 0) {
        bool ok = n != 0 && (((uintptr_t)out & 15) == 0);
        for (int i = 0; ok && i < 5; i++) {
            cudaPointerAttributes at;
            ok = cudaPointerGetAttributes(&at, outs[i]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer == outs[i];
        }
        cudaGetLastError();
        memcpy(s.zc_seen, outs, sizeof outs); s.zc_ok = ok;
    }
    s.zero_copy = s.zc_ok && n != 0;
    if (s.zero_copy) {
        const uint32_t zcap = out_cap < e->out_dev_cap ? out_cap : e->out_dev_cap;    // the bounce buffer bounds it too
        int rc = launch(e, s.d_desc, s.d_pkts, s.d_off, n, seed, qidx_base, out, zcap, out_off, out_len, status, miss_idx, s.h_totals, s.stream, s.d_out, flags);
        if (rc != BB_OK) return rc;
        CK(cudaEventRecord(s.ev, s.stream));
        s.busy = true; s.n = n; s.epoch = (uint32_t)e->epoch; s.out = out; s.out_cap = out_cap; s.miss_idx = miss_idx; s.n_miss = n_miss;
        return BB_OK;
    }
    uint32_t cap = out_cap < e->out_dev_cap ? out_cap : e->out_dev_cap;
    int rc = launch(e, s.d_desc, s.d_pkts, s.d_off, n, seed, qidx_base, s.d_out, cap, s.d_out_off, s.d_out_len, s.d_status, s.d_miss, s.d_totals, s.stream, nullptr, flags);
    if (rc != BB_OK) return rc;
    CK(cudaMemcpyAsync(s.h_totals, s.d_totals, 16, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaEventRecord(s.ev, s.stream));
    CK(cudaMemcpyAsync(out_off, s.d_out_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s.stream));
    if (n) CK(cudaMemcpyAsync(status, s.d_status, n, cudaMemcpyDeviceToHost, s.stream));
    if (n) CK(cudaMemcpyAsync(out_len, s.d_out_len, (size_t)n * 2, cudaMemcpyDeviceToHost, s.stream));
    s.busy = true; s.n = n; s.epoch = (uint32_t)e->epoch; s.out = out; s.out_cap = out_cap; s.miss_idx = miss_idx; s.n_miss = n_miss;
    return BB_OK;
}

int bb_resolve_submit(bb_engine* e, int slot, const uint8_t* pkts, const uint32_t* pkt_off, uint32_t n, uint64_t seed,
                      uint32_t qidx_base, uint8_t* out, uint32_t out_cap, uint32_t* out_off, uint16_t* out_len, uint8_t* status,
                      uint32_t* miss_idx, uint32_t* n_miss) {
    return bb_resolve_submit_ex(e, slot, pkts, pkt_off, n, seed, qidx_base, out, out_cap, out_off, out_len, status, miss_idx, n_miss, 0);
}

int bb_resolve_wait(bb_engine* e, int slot) {
    if (!e || slot < 0 || slot >= NSLOTS) return BB_ERR_ARG;
    SlotCtx& s = e->slots[slot];
    if (!s.busy) return BB_ERR_ARG;
    s.busy = false;
    CK(cudaSetDevice(e->device));
    CK(cudaEventSynchronize(s.ev));
    const uint32_t total = s.n ? s.h_totals[0] : 0, nmiss = s.n ? s.h_totals[1] : 0;
    const bool ovf = s.n && s.h_totals[2] == s.epoch;
    if (ovf || total > s.out_cap) { cudaStreamSynchronize(s.stream); return BB_ERR_CAPACITY; }
    if (s.zero_copy) { *s.n_miss = nmiss; return BB_OK; }        // everything is already in the caller's buffers
    if (total) CK(cudaMemcpyAsync(s.out, s.d_out, total, cudaMemcpyDeviceToHost, s.stream));
    if (nmiss) CK(cudaMemcpyAsync(s.miss_idx, s.d_miss, (size_t)nmiss * 4, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    *s.n_miss = nmiss;
    return BB_OK;
}

int bb_resolve_batch_ex(bb_engine* e, const uint8_t* pkts, const uint32_t* pkt_off, uint32_t n, uint64_t seed,
                        uint32_t qidx_base, uint8_t* out, uint32_t out_cap, uint32_t* out_off, uint16_t* out_len, uint8_t* status,
                        uint32_t* miss_idx, uint32_t* n_miss, uint32_t flags) {
    int rc = bb_resolve_submit_ex(e, 0, pkts, pkt_off, n, seed, qidx_base, out, out_cap, out_off, out_len, status, miss_idx, n_miss, flags);
    if (rc != BB_OK) return rc;
    return bb_resolve_wait(e, 0);
}
int bb_resolve_batch(bb_engine* e, const uint8_t* pkts, const uint32_t* pkt_off, uint32_t n, uint64_t seed,
                     uint32_t qidx_base, uint8_t* out, uint32_t out_cap, uint32_t* out_off, uint16_t* out_len, uint8_t* status,
                     uint32_t* miss_idx, uint32_t* n_miss) {
    return bb_resolve_batch_ex(e, pkts, pkt_off, n, seed, qidx_base, out, out_cap, out_off, out_len, status, miss_idx, n_miss, 0);
}

}  // extern "C"

// =================================================================================================
// multi-GPU shard: receive regions in peer-mapped HBM, route + push, region resolves
// =================================================================================================
struct bb_shard {
    bb_engine* e = nullptr; uint32_t nranks = 1, rank = 0, max_batch = 0, cap_q = 0, cap_b = 0;
    size_t reg_size = 0;
    // 2 x nranks regions: set (step & 1), region s = queries pushed by rank s.  Double-buffered so that
    // a fast rank's push of step k+1 never lands in a region a slower owner is still resolving (step k);
    // by the time step k+2 reuses the set, every rank has passed barrier k+1, i.e. finished resolve k.
    uint8_t* recv = nullptr;
    uint8_t* peer_recv[bbk::MAX_RANKS] = {};       // rank d's receive buffer as mapped here
    unsigned long long* cursor = nullptr; uint32_t* done = nullptr; uint32_t* err = nullptr;
    uint32_t epoch = 0;
    // owner side: one output set per source region, each kind contiguous with a fixed stride so that
    // all regions resolve in ONE launch (grid.y = region)
    uint8_t* d_out = nullptr; uint8_t* d_out_off = nullptr; uint8_t* d_out_len = nullptr; uint8_t* d_status = nullptr;
    uint8_t* d_miss = nullptr; uint8_t* d_totals = nullptr; uint8_t* d_desc = nullptr;
    size_t out_stride = 0, off_stride = 0, len_stride = 0, status_stride = 0, miss_stride = 0, totals_stride = 32, desc_stride = 0;
    uint32_t out_cap = 0;
    // optional pinned host mirrors of the result set (same strides): the resolve kernel writes them directly
    uint8_t* h_out = nullptr; uint8_t* h_out_off = nullptr; uint8_t* h_out_len = nullptr; uint8_t* h_status = nullptr;
    uint8_t* h_miss = nullptr; uint8_t* h_totals = nullptr; uint8_t* h_qidx = nullptr;
    bool host = false; uint32_t resolve_epoch = 0;
};

extern "C" {

bb_shard* bb_shard_create(bb_engine* e, uint32_t nranks, uint32_t rank, uint32_t max_batch, uint32_t bytes_per_query, int* err) {
    auto fail = [&](int c) -> bb_shard* { if (err) *err = c; return nullptr; };
    if (err) *err = BB_OK;
    if (!e || nranks == 0 || nranks > bbk::MAX_RANKS || rank >= nranks || max_batch == 0 || max_batch > e->max_batch) return fail(BB_ERR_ARG);
    bb_shard* s = new bb_shard();
    s->e = e; s->nranks = nranks; s->rank = rank; s->max_batch = max_batch;
    // A region must be able to hold a whole ingress batch: queries that need no lookup stay on the
    // ingress rank, and real DNS traffic is skewed (one hot name sends a full batch to one owner).
    s->cap_q = max_batch;
    s->cap_b = s->cap_q * (bytes_per_query ? bytes_per_query : 64);
    s->reg_size = bbk::region_size(s->cap_q, s->cap_b);
    s->out_cap = s->cap_q * 256u;
    auto ck = [&](cudaError_t c) { if (c != cudaSuccess) { g_cuda_err = cudaGetErrorString(c); return false; } return true; };
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    s->out_stride = up((size_t)s->out_cap + 64); s->off_stride = up(((size_t)s->cap_q + 1) * 4); s->len_stride = up((size_t)s->cap_q * 2 + 16);
    s->status_stride = up((size_t)s->cap_q + 16); s->miss_stride = up((size_t)s->cap_q * 4 + 16); s->desc_stride = up(((size_t)e->max_tiles + 4) * 8);
    bool ok = ck(cudaSetDevice(e->device)) && ck(cudaMalloc(&s->recv, s->reg_size * nranks * 2)) && ck(cudaMemset(s->recv, 0, s->reg_size * nranks * 2)) &&
              ck(cudaMalloc(&s->cursor, 8 * bbk::MAX_RANKS)) && ck(cudaMemset(s->cursor, 0, 8 * bbk::MAX_RANKS)) &&
              ck(cudaMalloc(&s->done, 16)) && ck(cudaMemset(s->done, 0, 16)) &&
              ck(cudaMalloc(&s->d_out, s->out_stride * nranks)) && ck(cudaMalloc(&s->d_out_off, s->off_stride * nranks)) &&
              ck(cudaMalloc(&s->d_out_len, s->len_stride * nranks)) && ck(cudaMalloc(&s->d_status, s->status_stride * nranks)) &&
              ck(cudaMalloc(&s->d_miss, s->miss_stride * nranks)) && ck(cudaMalloc(&s->d_totals, s->totals_stride * nranks)) &&
              ck(cudaMemset(s->d_totals, 0, s->totals_stride * nranks)) &&
              ck(cudaMalloc(&s->d_desc, s->desc_stride * nranks)) && ck(cudaMemset(s->d_desc, 0, s->desc_stride * nranks));
    s->err = s->done ? s->done + 2 : nullptr;
    if (!ok) { bb_shard_destroy(s); return fail(BB_ERR_CUDA); }
    s->peer_recv[rank] = s->recv;
    return s;
}

void bb_shard_destroy(bb_shard* s) {
    if (!s) return;
    cudaSetDevice(s->e->device); cudaDeviceSynchronize();
    for (uint32_t r = 0; r < s->nranks; r++) if (r != s->rank && s->peer_recv[r]) cudaIpcCloseMemHandle(s->peer_recv[r]);
    cudaFree(s->d_out); cudaFree(s->d_out_off); cudaFree(s->d_out_len); cudaFree(s->d_status); cudaFree(s->d_miss);
    cudaFree(s->d_totals); cudaFree(s->d_desc);
    cudaFree(s->recv); cudaFree(s->cursor); cudaFree(s->done);
    cudaFreeHost(s->h_out); cudaFreeHost(s->h_out_off); cudaFreeHost(s->h_out_len); cudaFreeHost(s->h_status);
    cudaFreeHost(s->h_miss); cudaFreeHost(s->h_totals); cudaFreeHost(s->h_qidx);
    delete s;
}

uint32_t bb_shard_ipc_handle_size(void) { return (uint32_t)sizeof(cudaIpcMemHandle_t); }
uint32_t bb_shard_region_capacity(const bb_shard* s) { return s ? s->cap_q : 0; }

int bb_shard_get_ipc_handle(bb_shard* s, void* out) {
    if (!s || !out) return BB_ERR_ARG;
    CK(cudaSetDevice(s->e->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->recv));
    memcpy(out, &h, sizeof h);
    return BB_OK;
}

// handles: nranks x bb_shard_ipc_handle_size() bytes, rank-major (this rank's own entry is ignored)
int bb_shard_open_peers(bb_shard* s, const void* handles) {
    if (!s || !handles) return BB_ERR_ARG;
    CK(cudaSetDevice(s->e->device));
    for (uint32_t r = 0; r < s->nranks; r++) {
        if (r == s->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * sizeof h, sizeof h);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->peer_recv[r] = (uint8_t*)p;
    }
    return BB_OK;
}

// Ingress: route every query of a device-resident batch to its owner and push it there.
int bb_shard_route_push(bb_shard* s, const uint8_t* d_pkts, const uint32_t* d_pkt_off, uint32_t n, uint32_t qidx_base, void* stream) {
    if (!s || n > s->max_batch || ((uintptr_t)d_pkts & 15)) return BB_ERR_ARG;
    for (uint32_t r = 0; r < s->nranks; r++) if (!s->peer_recv[r]) return BB_ERR_ARG;
    bb_engine* e = s->e;
    bbk::PushParams A; memset(&A, 0, sizeof A);
    A.P.pkts = d_pkts; A.P.pkt_off = d_pkt_off; A.P.n = n; A.P.eng = e->d_const; A.P.ready = 1;
    A.P.suffix_len = e->hconst.suffix_len; A.P.soa_len = e->hconst.soa_len; A.P.recursion = e->hconst.recursion;
    A.P.route = 1; A.P.nranks = s->nranks; A.P.rank = s->rank; A.P.table = e->d_table; A.P.mask = e->mask; A.P.arena = e->d_arena;
    A.epoch = ++s->epoch;
    const size_t set = (size_t)(s->epoch & 1) * s->nranks;
    for (uint32_t r = 0; r < s->nranks; r++) A.region[r] = s->peer_recv[r] + (set + s->rank) * s->reg_size;
    A.cap_q = s->cap_q; A.cap_b = s->cap_b; A.cursor = s->cursor; A.done = s->done; A.err = s->err;
    A.qidx_base = qidx_base; A.P.stage_log = e->stage_log;
    // n == 0 still publishes empty region headers (one block, no queries)
    const uint32_t grid = n ? (n + bbk::T - 1) / bbk::T : 1;
    bbk::route_push_kernel<<<grid, bbk::T, 0, (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError());
    e->launches++;
    return BB_OK;
}

// Owner: resolve the nranks receive regions (after the caller's cross-rank barrier), one launch
// per region on forked streams that join back into `stream`.
int bb_shard_resolve(bb_shard* s, uint64_t seed, int wait_for_peers, void* stream) {
    if (!s) return BB_ERR_ARG;
    bb_engine* e = s->e;
    cudaStream_t main = (cudaStream_t)stream;
    if (wait_for_peers) {
        bbk::wait_regions_kernel<<<1, 32, 0, main>>>(s->recv + (size_t)(s->epoch & 1) * s->nranks * s->reg_size, s->reg_size, s->nranks,
                                                      s->epoch, s->err);
        CK(cudaGetLastError());
    }
    uint8_t* reg = s->recv + (size_t)(s->epoch & 1) * s->nranks * s->reg_size;          // region 0 of the set just pushed
    bbk::Params P; memset(&P, 0, sizeof P);
    P.pkts = reg + bbk::region_bytes(s->cap_q); P.pkt_off = (const uint32_t*)(reg + bbk::region_off_array(s->cap_q));
    P.n = 0; P.n_dev = (const uint32_t*)reg; P.qidx_map = (const uint32_t*)(reg + bbk::region_qidx_array(s->cap_q));
    P.seed = seed; P.qidx_base = 0;
    P.out = s->d_out; P.out_cap = s->out_cap; P.out_off = (uint32_t*)s->d_out_off; P.out_len = (uint16_t*)s->d_out_len;
    P.status = s->d_status; P.miss_idx = (uint32_t*)s->d_miss; P.totals = (uint32_t*)s->d_totals;
    P.err_in = s->err;
    if (s->host) {                                   // results straight into the pinned mirrors (zero-copy)
        P.out = s->h_out; P.out_off = (uint32_t*)s->h_out_off; P.out_len = (uint16_t*)s->h_out_len; P.status = s->h_status;
        P.miss_idx = (uint32_t*)s->h_miss; P.totals = (uint32_t*)s->h_totals; P.qidx_out = (uint32_t*)s->h_qidx;
        P.bounce = s->d_out;                         // tiles too large to stage assemble on the device first
    }
    P.table = e->d_table; P.mask = e->mask; P.arena = e->d_arena; P.ready = e->ready && e->d_table; P.eng = e->d_const;
    P.suffix_len = e->hconst.suffix_len; P.soa_len = e->hconst.soa_len; P.recursion = e->hconst.recursion;
    P.ntiles = (s->cap_q + bbk::T - 1) / bbk::T; P.ntiles_cap = e->max_tiles;
    P.desc = (unsigned long long*)s->d_desc; P.counter = (uint32_t*)((unsigned long long*)s->d_desc + e->max_tiles + 1);
    P.epoch = (uint32_t)(++e->epoch); P.stage_log = nullptr; P.route = 0; P.nranks = s->nranks; P.rank = s->rank;
    P.regions = 1; P.in_stride = s->reg_size; P.out_stride = s->out_stride; P.off_stride = s->off_stride; P.len_stride = s->len_stride;
    P.status_stride = s->status_stride; P.miss_stride = s->miss_stride; P.totals_stride = s->totals_stride; P.desc_stride = s->desc_stride;
    const dim3 grid(P.ntiles, s->nranks);
    if (e->ordered) bbk::resolve_kernel<true, true><<<grid, bbk::T, 0, main>>>(P);
    else bbk::resolve_kernel<false, true><<<grid, bbk::T, 0, main>>>(P);
    CK(cudaGetLastError());
    s->resolve_epoch = P.epoch;
    e->launches++;
    return BB_OK;
}

// Pinned host mirrors of the per-region result set: from now on bb_shard_resolve writes responses,
// offsets, lengths, statuses, ingress indices, miss lists and totals straight into host memory
// (no copies, no size round trip), and bb_shard_results hands out pointers into them.
int bb_shard_host_results(bb_shard* s, int enable) {
    if (!s) return BB_ERR_ARG;
    CK(cudaSetDevice(s->e->device));
    if (enable && !s->h_out) {
        const unsigned fl = cudaHostAllocPortable | cudaHostAllocMapped;
        const size_t R = s->nranks;
        CK(cudaHostAlloc((void**)&s->h_out, s->out_stride * R, fl)); CK(cudaHostAlloc((void**)&s->h_out_off, s->off_stride * R, fl));
        CK(cudaHostAlloc((void**)&s->h_out_len, s->len_stride * R, fl)); CK(cudaHostAlloc((void**)&s->h_status, s->status_stride * R, fl));
        CK(cudaHostAlloc((void**)&s->h_miss, s->miss_stride * R, fl)); CK(cudaHostAlloc((void**)&s->h_totals, s->totals_stride * R, fl));
        CK(cudaHostAlloc((void**)&s->h_qidx, s->off_stride * R, fl));
        memset(s->h_totals, 0, s->totals_stride * R);
    }
    CK(cudaDeviceSynchronize());                     // resolves in flight finish with the old destination
    s->host = enable != 0;
    return BB_OK;
}

// Region `src` of the last bb_shard_resolve, in the host mirrors.  The caller has waited for the work
// it enqueued on the resolve's stream (event or stream synchronize); BB_ERR_ARG if that resolve has
// not finished writing this region.  The pointers stay valid until the next resolve on this shard.
int bb_shard_results(bb_shard* s, uint32_t src, const uint8_t** out, const uint32_t** out_off, const uint16_t** out_len,
                     const uint8_t** status, const uint32_t** qidx, const uint32_t** miss_idx,
                     uint32_t* n_out, uint32_t* n_miss, uint32_t* total_out) {
    if (!s || src >= s->nranks || !s->host || !s->resolve_epoch) return BB_ERR_ARG;
    const volatile uint32_t* tot = (const volatile uint32_t*)(s->h_totals + src * s->totals_stride);
    if (tot[3] != s->resolve_epoch) return BB_ERR_ARG;
    if (tot[6] == 2) { g_cuda_err = "timed out waiting for a peer rank's push"; return BB_ERR_CUDA; }
    if (tot[5] || tot[2] == s->resolve_epoch) return BB_ERR_CAPACITY;      // a sender overflowed the region / responses overflowed out
    const uint32_t n = tot[4];
    *n_out = n; *n_miss = n ? tot[1] : 0; *total_out = n ? tot[0] : 0;
    *out = s->h_out + src * s->out_stride; *out_off = (const uint32_t*)(s->h_out_off + src * s->off_stride);
    *out_len = (const uint16_t*)(s->h_out_len + src * s->len_stride); *status = s->h_status + src * s->status_stride;
    *qidx = (const uint32_t*)(s->h_qidx + src * s->off_stride); *miss_idx = (const uint32_t*)(s->h_miss + src * s->miss_stride);
    return BB_OK;
}

// Results of region `src` to host memory (synchronous).  qidx[i] is the index the query had in
// rank src's ingress numbering (qidx_base + position).  Returns the region's query count in *n_out.
int bb_shard_fetch(bb_shard* s, uint32_t src, uint8_t* out, uint32_t out_cap, uint32_t* out_off, uint16_t* out_len,
                   uint8_t* status, uint32_t* qidx, uint32_t* miss_idx, uint32_t* n_out, uint32_t* n_miss, uint32_t* total_out) {
    if (!s || src >= s->nranks || s->host) return BB_ERR_ARG;       // host mirrors on: use bb_shard_results
    CK(cudaSetDevice(s->e->device));
    uint8_t* reg = s->recv + ((size_t)(s->epoch & 1) * s->nranks + src) * s->reg_size;
    uint32_t hdr[4], tot[4];
    CK(cudaMemcpy(hdr, reg, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tot, s->d_totals + src * s->totals_stride, 16, cudaMemcpyDeviceToHost));
    if (hdr[3]) return BB_ERR_CAPACITY;                          // a sender overflowed this region
    { uint32_t werr = 0; CK(cudaMemcpy(&werr, s->err, 4, cudaMemcpyDeviceToHost)); if (werr == 2) { g_cuda_err = "timed out waiting for a peer rank's push"; return BB_ERR_CUDA; } }
    const uint32_t n = hdr[0];
    *n_out = n; *n_miss = n ? tot[1] : 0; *total_out = n ? tot[0] : 0;
    if (!n) return BB_OK;
    if (tot[0] > out_cap) return BB_ERR_CAPACITY;
    CK(cudaMemcpy(out, s->d_out + src * s->out_stride, tot[0], cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_off, s->d_out_off + src * s->off_stride, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out_len, s->d_out_len + src * s->len_stride, (size_t)n * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(status, s->d_status + src * s->status_stride, n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(qidx, reg + bbk::region_qidx_array(s->cap_q), (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (tot[1]) CK(cudaMemcpy(miss_idx, s->d_miss + src * s->miss_stride, (size_t)tot[1] * 4, cudaMemcpyDeviceToHost));
    return BB_OK;
}

}  // extern "C"
