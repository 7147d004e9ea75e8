// tests/dropin/caller_test.cpp -- TEST: include/pf_caller.hpp (the batched per-bubble caller over the C ABI) fed with the bubbles
// the unmodified reference aligned in tests/golden/e2e (their raw branch strings, entrance / exit ids and sizes) must write the
// reference's own files.  usage: caller_test <fixture dir> <out dir> <lower> <upper> [mt|t1] [host threads]; `mt` = the `-t N` dialect.  tests/test_gpu_dropin.py builds it with g++
// against libpfgpu.so, runs it and compares the files byte for byte.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "pf_caller.hpp"

static std::vector<std::string> split_tab(const std::string &s) {
    std::vector<std::string> out;
    std::string cur;
    for (char c : s) { if (c == '\t') { out.push_back(cur); cur.clear(); } else cur += c; }
    out.push_back(cur);
    return out;
}

static void write_file(const std::string &path, const std::string &text) {
    std::ofstream f(path, std::ios::binary);
    f << text;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: caller_test <fixture dir> <out dir> <lower> <upper>\n"); return 2; }
    const std::string fx = argv[1], od = argv[2];
    const unsigned lower = (unsigned)atoi(argv[3]), upper = (unsigned)atoi(argv[4]);
    std::map<unsigned, size_t> usize;
    {
        std::ifstream f(fx + "/P_Unitig_Id.txt");
        std::string ln;
        while (std::getline(f, ln)) { auto p = split_tab(ln); if (p.size() >= 2) usize[(unsigned)atoi(p[0].c_str())] = p[1].size(); }
    }
    std::vector<pfdropin::Bubble> batch;
    {
        std::ifstream f(fx + "/P_alignseq.txt");
        std::string ln, last_id;
        while (std::getline(f, ln)) {
            auto p = split_tab(ln);
            if (p.size() < 5) continue;
            if (batch.empty() || p[0] != last_id) {
                pfdropin::Bubble b;
                b.strict = p[1] == "1";
                b.entrance_id = (unsigned)atoi(p[2].c_str()); b.exit_id = (unsigned)atoi(p[3].c_str());
                b.entrance_size = usize[b.entrance_id]; b.exit_size = usize[b.exit_id];
                batch.push_back(b);
                last_id = p[0];
            }
            std::string raw;
            for (char c : p[4]) if (c != '-') raw += c;
            batch.back().branches.push_back(raw);
        }
    }
    pf_ctx *ctx = nullptr;
    pf_kmc *db = nullptr;
    if (pf_init(0, &ctx) != PF_OK) { fprintf(stderr, "pf_init: %s\n", pf_last_error()); return 1; }
    if (pf_kmc_open(ctx, (fx + "/db").c_str(), &db) != PF_OK) { fprintf(stderr, "pf_kmc_open: %s\n", pf_last_error()); return 1; }
    pfdropin::BubbleCaller caller(ctx, db, 2, -1, -3, lower, upper);
    const bool mt = argc > 5 && std::string(argv[5]) == "mt";
    caller.set_thread_dialect(mt);
    if (argc > 6) caller.set_host_threads((unsigned)atoi(argv[6]));
    pfdropin::CallerFiles out;
    size_t var_id = mt ? 0 : 1;
    // two calls: the caller is batched, the files must not depend on where the batches are cut
    std::vector<pfdropin::Bubble> first(batch.begin(), batch.begin() + batch.size() / 3), second(batch.begin() + batch.size() / 3, batch.end());
    if (!caller.call(first, var_id, out) || !caller.call(second, var_id, out)) { fprintf(stderr, "caller: %s\n", caller.error().c_str()); return 1; }
    static const char *names[4] = {"bi", "tri", "tetra", "penta"};
    write_file(od + "/P_alignseq.txt", out.alignseq);
    write_file(od + "/P_allele_frequency.txt", out.allele_frequency);
    for (int i = 0; i < 4; i++) {
        write_file(od + "/P_" + names[i] + "cov.txt", out.cov[i]);
        write_file(od + "/P_" + names[i] + "fre.txt", out.fre[i]);
    }
    printf("bubbles %zu called %zu alleles 2:%zu 3:%zu 4:%zu 5:%zu\n", batch.size(), out.bubbles_called, out.alleles[0], out.alleles[1],
           out.alleles[2], out.alleles[3]);
    pf_kmc_close(db);
    pf_shutdown(ctx);
    return 0;
}
