// Sanitizer fuzz of the host-side C code (no GPU): the balancer frame parser/builder under tight capacities and
// corrupted streams, the zone builder / delta applier on garbage, truncated and corrupted JSON lines.  Built with
// -fsanitize=address,undefined by tests/test_native_sanitize.py; iteration counts come from argv.
#include "binder_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
int main(int argc, char** argv) {
    const int n_frames = argc > 1 ? atoi(argv[1]) : 20000, n_zones = argc > 2 ? atoi(argv[2]) : 4000;
    std::mt19937 rng(7);
    // 1. frame parser on corrupted / random streams with tight capacities
    for (int it = 0; it < n_frames; it++) {
        std::vector<uint8_t> in;
        int nf = rng() % 8;
        for (int f = 0; f < nf; f++) {
            uint32_t t = (rng() % 10 < 7) ? 2u : (uint32_t[]){1u, 4u, 3u, 77u, 1002u}[rng() % 5];
            uint32_t len = rng() % 10 == 0 ? rng() % 3000 : rng() % 80;
            uint32_t w[4] = { t, (uint32_t)rng(), (uint32_t)rng() % 65536, len };
            const uint8_t* p = (const uint8_t*)w;
            if (t == 2) { in.insert(in.end(), p, p + 16); for (uint32_t i = 0; i < len; i++) in.push_back((uint8_t)rng()); }
            else in.insert(in.end(), p, p + 4);
        }
        if (rng() % 3 == 0 && !in.empty()) in.resize(rng() % in.size());          // truncated
        if (rng() % 5 == 0) for (auto& b : in) if (rng() % 50 == 0) b = (uint8_t)rng();
        uint32_t cap_n = rng() % 6, cap_b = rng() % 300, cap_c = rng() % 3;
        std::vector<uint8_t> pk(cap_b + 1); std::vector<uint32_t> off(cap_n + 1), ip(cap_n + 1), port(cap_n + 1), ctl(cap_c + 1);
        uint32_t n = 0, nc = 0; size_t used = 0;
        int rc = bb_frames_parse(in.data(), in.size(), pk.data(), cap_b, off.data(), ip.data(), port.data(), cap_n, &n, ctl.data(), cap_c, &nc, &used);
        if (used > in.size() || n > cap_n || nc > cap_c || off[n] > cap_b || (rc != 0 && rc != BB_ERR_PROTOCOL)) { printf("BAD parse\n"); return 1; }
        // build frames from fake results of that batch
        std::vector<uint16_t> rl(n + 1); std::vector<uint8_t> st(n + 1); std::vector<uint32_t> ro(n + 2, 0); std::vector<uint8_t> resp;
        for (uint32_t i = 0; i < n; i++) { st[i] = rng() % 3; rl[i] = st[i] ? 0 : rng() % 100; ro[i] = (uint32_t)resp.size(); resp.resize(resp.size() + rl[i], 0xAB); }
        size_t need = 0; bb_frames_build(resp.data(), ro.data(), rl.data(), st.data(), ip.data(), port.data(), n, ctl.data(), nc, nullptr, 0, &need);
        std::vector<uint8_t> out(need + 1);
        size_t got = 0; int r2 = bb_frames_build(resp.data(), ro.data(), rl.data(), st.data(), ip.data(), port.data(), n, ctl.data(), nc, out.data(), need, &got);
        if (r2 != BB_OK || got != need) { printf("BAD build\n"); return 1; }
    }
    // 2. zone builder / delta parser on garbage lines
    const char* frag[] = { "{\"path\":\"/com/foo/a\",\"data\":{\"type\":\"host\",\"host\":{\"address\":\"1.2.3.4\"}}}", "{\"path\":\"/com/foo\",\"data\":null}",
        "{\"path\":\"/com/foo/a\",\"deleted\":true}", "{\"path\":\"/com/foo/a/b/c\",\"raw\":\"{\\\"type\\\":\\\"service\\\",\\\"service\\\":{}}\"}", "{", "}", "[]", "null", "\"x\"",
        "{\"path\":5}", "{\"path\":\"/com/foo/\\ud800x\",\"data\":{}}", "{\"path\":\"/com/foo//x\",\"data\":[]}", "{\"path\":\"/com/foo/s\",\"data\":{\"type\":\"service\",\"service\":{\"service\":null}}}",
        "{\"path\":\"/com/foo/s/k\",\"data\":{\"type\":\"rr_host\",\"rr_host\":{\"address\":\"1.1.1.1\",\"ports\":[1,2,70000,-1,1e400]},\"ttl\":1e99}}" };
    for (int it = 0; it < n_zones; it++) {
        std::string snap, delta;
        int nl = rng() % 12;
        for (int i = 0; i < nl; i++) { std::string l = frag[rng() % (sizeof frag / sizeof *frag)]; if (rng() % 6 == 0 && !l.empty()) l.resize(rng() % l.size()); if (rng() % 8 == 0) for (auto& c : l) if (rng() % 20 == 0) c = (char)rng(); (rng() % 2 ? snap : delta) += l + "\n"; }
        int err = 0; bb_zone* z = bb_zone_build(snap.data(), snap.size(), "foo.com", &err);
        if (!z) { z = bb_zone_build("", 0, "foo.com", &err); }
        if (z) { bb_zone_apply(z, delta.data(), delta.size()); bb_zone_apply(z, snap.data(), snap.size()); uint8_t k; uint32_t a, b, c; bb_zone_probe(z, 0, (const uint8_t*)"a.foo.com", 9, &k, &a, &b, nullptr, 0, &c); bb_zone_free(z); }
    }
    printf("OK\n");
    return 0;
}
