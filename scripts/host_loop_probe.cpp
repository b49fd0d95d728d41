// Host-side timing of the bsdiff greedy loop (scan / extend / write) over a (pos, len) table dumped to disk.
//   g++ -O2 -pthread -o /tmp/probe scripts/host_loop_probe.cpp && /tmp/probe DIR   (DIR holds old.bin new.bin pos.bin len.bin)
#include "../deltaq_b200/csrc/dq_diff_host.h"
#include <chrono>
#include <cstdio>
#include <fstream>
#include <string>
using namespace dq::diffhost;
template <class T> std::vector<T> rd(const std::string &p)
{
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    size_t n = f.tellg();
    f.seekg(0);
    std::vector<T> v(n / sizeof(T));
    f.read((char *)v.data(), n);
    return v;
}
double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    std::string d = argc > 1 ? argv[1] : "/tmp/hl";
    auto old = rd<uint8_t>(d + "/old.bin"), nw = rd<uint8_t>(d + "/new.bin");
    auto pos = rd<int32_t>(d + "/pos.bin"), len = rd<int32_t>(d + "/len.bin");
    int n = old.size(), m = nw.size();
    // host mirror of encode_table_kernel
    std::vector<uint8_t> code(m + 64);
    std::vector<MatchHead> heads;
    std::vector<TileEntry> tiles((m + 1023) / 1024);
    for (int t = 0; t < (int)tiles.size(); ++t) {
        tiles[t].base = heads.size();
        for (int x = t * 1024; x < std::min(m, (t + 1) * 1024); ++x) {
            code[x] = len[x] < 9 ? len[x] : 9;
            bool cont = x > 0 && pos[x] == pos[x - 1] + 1 && len[x] == len[x - 1] - 1;
            if (len[x] >= 9 && !cont) heads.push_back({x, pos[x], len[x]});
        }
        tiles[t].count = heads.size() - tiles[t].base;
    }
    printf("n %d m %d heads %zu\n", n, m, heads.size());
    {
        auto fetch = [&](int s) { return pos[s]; };
        double best = 1e9;
        for (int r = 0; r < 30; ++r) {
            CodedTable<decltype(fetch)> ct{code.data(), tiles.data(), heads.data(), (uint32_t)heads.size(), fetch};
            Streams tmp;
            int cnt = 0;
            double t = now();
            greedy_scan(old.data(), n, nw.data(), m, ct, tmp, [](int) {}, [&](int, int) { ++cnt; });
            best = std::min(best, now() - t);
        }
        printf("coded scan, best of 30: %.2f ms\n", best);
    }
    Streams out, o2, o3;
    for (int rep = 0; rep < 4; rep++) {
        reset_streams(out, m);
        struct Stop { int scan, pos; };
        std::vector<Stop> stops;
        FullTable tab{pos.data(), len.data()};
        double t0 = now();
        greedy_scan(old.data(), n, nw.data(), m, tab, out, [](int) {}, [&](int s, int p) { stops.push_back({s, p}); });
        double t1 = now();
        auto fetch = [&](int s) { return pos[s]; };
        CodedTable<decltype(fetch)> ct{code.data(), tiles.data(), heads.data(), (uint32_t)heads.size(), fetch};
        std::vector<Stop> stops2;
        Streams tmp;
        greedy_scan(old.data(), n, nw.data(), m, ct, tmp, [](int) {}, [&](int s, int p) { stops2.push_back({s, p}); });
        double t2 = now();
        bool same = stops.size() == stops2.size() && tmp.visits == out.visits;
        for (size_t i = 0; same && i < stops.size(); ++i) same = stops[i].scan == stops2[i].scan && stops[i].pos == stops2[i].pos;
        EmitState st;
        std::vector<Piece> pcs;
        for (auto &s : stops) pcs.push_back(extend_stop(old.data(), n, nw.data(), m, s.scan, s.pos, st));
        double t3 = now();
        {
            // the same extensions with the stretches the scan certifies (exact table): identical pieces, less walking
            std::vector<std::vector<Cert>> per_piece(1);
            long long cert_bytes = 0;
            Streams tmp2;
            CodedTable<decltype(fetch)> ct3{code.data(), tiles.data(), heads.data(), (uint32_t)heads.size(), fetch};
            greedy_scan(old.data(), n, nw.data(), m, ct3, tmp2, [](int) {},
                        [&](int, int) { per_piece.emplace_back(); },
                        [&](int s, int l) { per_piece.back().push_back(Cert{s, l}); cert_bytes += l; });
            double ta = now();
            EmitState st2;
            bool same_pieces = true;
            for (size_t k = 0; k < stops.size(); ++k) {
                Piece pc = extend_stop(old.data(), n, nw.data(), m, stops[k].scan, stops[k].pos, st2, nullptr,
                                       per_piece[k].data(), per_piece[k].size());
                same_pieces = same_pieces && pc.lastscan == pcs[k].lastscan && pc.lastpos == pcs[k].lastpos &&
                              pc.lenf == pcs[k].lenf && pc.extra == pcs[k].extra && pc.seek == pcs[k].seek;
            }
            printf("extend with certified stretches %.2f ms (%s), %lld of %d bytes certified\n", now() - ta,
                   same_pieces ? "same pieces" : "DIFFERENT", cert_bytes, m);
        }
        const double t3b = now();
        for (auto &pc : pcs) write_piece(old.data(), nw.data(), pc, out);
        double t4 = now() - (t3b - t3);
        CodedTable<decltype(fetch)> ct2{code.data(), tiles.data(), heads.data(), (uint32_t)heads.size(), fetch};
        greedy_emit_pipelined(old.data(), n, nw.data(), m, ct2, o2, [](int) {});
        double t5 = now();
        printf("scan full %.2f ms | scan coded %.2f ms (%s) | extend %.2f | write %.2f | pipelined coded %.2f ms | %zu stops, visits %lld\n",
               t1 - t0, t2 - t1, same ? "same stops" : "DIFFERENT", t3 - t2, t4 - t3, t5 - t4, stops.size(), (long long)out.visits);
    }
}
