// integration/ploidy_estimation_colored_gpu.cpp -- the reference-side binding of the COLOURED mode (`-f <graph>.bfg_colors`, one KMC
// database and one (low, up) gate per colour; BASELINE configs[3]).
//
// Replacement DEFINITIONS of CCDBG::ploidyEstimation_ptr (`-t 1`, declared src/CCDBG.hpp:35, defined src/CCDBG.cpp:2759) and
// CCDBG::ploidyEstimation_multithread_ptr (`-t N`, CCDBG.hpp:24, CCDBG.cpp:538): the per-superbubble analysis -- readCovUni of the
// entrance and of the branch unitigs per colour, SeqAlign, site k-mers, readCov per colour, Cramer's V, output rows -- handed to
// libpfgpu.so in flat batches through pfdropin::ColoredBubbleCaller (include/pf_caller_colored.hpp).  What stays here is what needs
// the graph: the walk over the unitigs with the reference's visited marks, the colour sets of the branch unitigs
// (UnitigColors::contains / size, CCDBG.cpp:699, :714) and of the unitigs that hold the site k-mers (findUnitig, :1127).
// integration/Makefile weakens the two symbols in the reference's CCDBG object and links this file next to the single-sample
// binding into `PloidyFrost_gpu`.
//
// Organisation as in ploidy_estimation_gpu.cpp: blocks of unitigs are SPECULATED by T host threads (everything about a possible
// bubble that does not depend on other bubbles, written into per-thread flat arenas), RESOLVED in graph order against the visited
// marks (the only coupling between bubbles), and handed to the device thread while the next block is speculated.
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "CCDBG.hpp"               // the reference's class (Bifrost coloured graph, MyUnitig marks)
#include "SeqAlign.hpp"            // the reference's own aligner: takes the bubbles that exceed a device limit
#include "pf_caller_colored.hpp"   // ours

using namespace std;

namespace {

typedef UnitigColorMap<MyUnitig> UCM;

vector<string> args_of_this_process() {
    ifstream f("/proc/self/cmdline", ios::binary);
    vector<string> args;
    string cur;
    char c;
    while (f.get(c)) { if (c == '\0') { args.push_back(cur); cur.clear(); } else cur += c; }
    return args;
}
// Main.cpp's getopt loop: `-x <value>` or `-x<value>`, the last occurrence wins
string option_of_this_process(char flag) {
    const vector<string> args = args_of_this_process();
    string val;
    for (size_t i = 1; i < args.size(); i++) {
        const string &a = args[i];
        if (a.size() >= 2 && a[0] == '-' && a[1] == flag) {
            if (a.size() > 2) val = a.substr(2);
            else if (i + 1 < args.size()) val = args[++i];
        }
    }
    return val;
}

// The reference opens one CKMCFile per line of the `-d` file in CCDBG's constructor (CCDBG.cpp:12-85) and keeps no names; the
// list is read again here.  Context + databases come up on a second thread from program start (only for coloured command lines)
// and the estimation phase joins it.
struct ColoredDevice {
    thread worker;
    pf_ctx *ctx = nullptr;
    vector<pf_kmc *> dbs;
    string error;

    void open_now(const string &list_file) {
        ifstream in(list_file);
        if (in.fail()) { error = "cannot read the kmc database name file " + list_file; return; }
        if (pf_init(0, &ctx) != PF_OK) { error = pf_last_error(); ctx = nullptr; return; }
        string name;
        while (getline(in, name, '\n')) {
            pf_kmc *db = nullptr;
            if (pf_kmc_open(ctx, name.c_str(), &db) != PF_OK) { error = pf_last_error(); return; }
            dbs.push_back(db);
        }
    }
    ColoredDevice() {
        const string list_file = option_of_this_process('d');
        if (!list_file.empty() && !option_of_this_process('g').empty() && !option_of_this_process('f').empty())
            worker = thread([this, list_file] { open_now(list_file); });
    }
    void ready() {
        if (worker.joinable()) worker.join();
        else if (!ctx && error.empty()) open_now(option_of_this_process('d'));
    }
    void release() {
        for (pf_kmc *db : dbs) pf_kmc_close(db);
        dbs.clear();
        if (ctx) pf_shutdown(ctx);
        ctx = nullptr;
    }
    ~ColoredDevice() {
        if (worker.joinable()) worker.join();
        release();
    }
};
ColoredDevice g_colored_device;

double seconds_since(const chrono::steady_clock::time_point &t0) {
    return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}

enum SpecKind : uint8_t { SPEC_COMPLEX = 0, SPEC_OTHER_END = 1, SPEC_BUBBLE = 2 };
struct Spec {
    MyUnitig *ud, *exit_ud;
    uint8_t strand, exit_strand, kind;
    uint32_t bubble;      // SPEC_BUBBLE: index in the thread's ColoredBatch
};
struct ThreadOut {
    vector<Spec> specs;
    pfdropin::ColoredBatch batch;
    void clear() { specs.clear(); batch.clear(); }
};

inline MyUnitig *data_of(const UCM &u) { return u.getData()->getData(u); }

// every path from the entrance to the exit, spelled from the entrance's last k-mer to the exit's first k-mer (CCDBG.cpp:936-981):
// depth first, a unitig's successors taken last-to-first (the order in which the reference's stack hands them out)
void spell_paths(const UCM &cur, const UCM &exit_uni, size_t k, string &text, pfdropin::ColoredBatch &out) {
    vector<UCM> next;
    for (const auto &nx : cur.getSuccessors()) next.push_back(nx);
    for (size_t i = next.size(); i-- > 0;) {
        const UCM &v = next[i];
        const string s = v.mappedSequenceToString();
        const size_t mark = text.size();
        if (v.isSameReferenceUnitig(exit_uni)) {
            text.push_back(s[k - 1]);
            out.add_branch(text.data(), text.size(), true, 0, false);
        } else {
            text.append(s, k - 1, string::npos);
            spell_paths(v, exit_uni, k, text, out);
        }
        text.resize(mark);
    }
}

void speculate(vector<UCM> &units, size_t i0, size_t i1, size_t k, size_t n_colors, ThreadOut &o) {
    o.clear();
    string text;
    for (size_t i = i0; i < i1; i++) {
        UCM u = units[i];
        MyUnitig *ud = data_of(u);
        for (int pass = 0; pass < 2; pass++) {
            const bool strand = pass == 0;
            if (strand ? ud->is_plus_visited() : ud->is_minus_visited()) continue;
            Spec sp;
            sp.ud = ud; sp.exit_ud = nullptr; sp.strand = strand; sp.exit_strand = 0; sp.bubble = 0;
            if (ud->isComplex(strand)) { sp.kind = SPEC_COMPLEX; o.specs.push_back(sp); continue; }
            u.strand = strand;
            const bool strict = ud->isStrict(strand);
            UCM exit_uni;
            if (strict) exit_uni = *u.getSuccessors().begin()->getSuccessors().begin();        // CCDBG.cpp:672
            else {                                                                             // :923-927
                const size_t want = ud->get_bubble_id(strand);
                exit_uni = *u.getSuccessors().begin();
                while (data_of(exit_uni)->get_id() != want) exit_uni = *exit_uni.getSuccessors().begin();
            }
            sp.exit_ud = data_of(exit_uni);
            sp.exit_strand = exit_uni.strand;
            const string ref = u.referenceUnitigToString();
            // reported from one end only: the end whose unitig string is not the smaller one (:674, :928); two different unitigs
            // differ inside their first k-mer, so the heads decide
            bool other_end;
            if (u.isSameReferenceUnitig(exit_uni)) other_end = false;
            else other_end = ref.compare(0, k, exit_uni.getUnitigHead().toString()) < 0;
            if (other_end) { sp.kind = SPEC_OTHER_END; o.specs.push_back(sp); continue; }
            sp.kind = SPEC_BUBBLE;
            sp.bubble = (uint32_t)o.batch.flat.n_bubbles();
            if (strict) {
                for (const auto &uu : u.getSuccessors()) {
                    const string s = uu.mappedSequenceToString();
                    const UnitigColors *uc = uu.getData()->getUnitigColors(uu);
                    uint64_t mask = 0;
                    size_t j = 0;
                    for (size_t c = 0; c < n_colors; c++)
                        if (uc->contains(uu, c)) { mask |= 1ull << c; j++; }                   // :699
                    o.batch.add_branch(s.data(), s.size(), uu.strand, mask, uc->size(uu) == j * uu.len);   // :714
                }
            } else {
                const string s = u.mappedSequenceToString();
                text.assign(s, s.size() - k, k);
                spell_paths(u, exit_uni, k, text, o.batch);
            }
            o.batch.end_bubble(strict, (unsigned)ud->get_id(), (unsigned)sp.exit_ud->get_id(), u.size, exit_uni.size, ref.data(), ref.size());
            o.specs.push_back(sp);
        }
    }
}

// the reference's marks, in its order (plus strand before minus, unitigs in graph order; CCDBG.cpp:2807-2835, :3504-3530)
void resolve(vector<ThreadOut> &outs, pfdropin::ColoredBatch &blk) {
    blk.clear();
    for (ThreadOut &o : outs)
        for (const Spec &sp : o.specs) {
            MyUnitig *ud = sp.ud;
            if (sp.strand ? ud->is_plus_visited() : ud->is_minus_visited()) continue;
            if (sp.strand) ud->set_plus_visited(); else ud->set_minus_visited();
            if (sp.kind != SPEC_BUBBLE) continue;
            blk.append_bubble(o.batch, sp.bubble);
            if (sp.exit_strand) sp.exit_ud->set_minus_visited(); else sp.exit_ud->set_plus_visited();
        }
}

struct Output {
    ofstream allfre, alignseq, cov[4], fre[4];
    bool open_all(const string &dir) {
        static const char *names[4] = {"bi", "tri", "tetra", "penta"};
        allfre.open(dir + "_allele_frequency.txt", ios::out | ios::trunc | ios::binary);
        alignseq.open(dir + "_alignseq.txt", ios::out | ios::trunc | ios::binary);
        bool ok = allfre.is_open() && alignseq.is_open();
        for (int i = 0; i < 4; i++) {
            cov[i].open(dir + "_" + names[i] + "cov.txt", ios::out | ios::trunc | ios::binary);
            fre[i].open(dir + "_" + names[i] + "fre.txt", ios::out | ios::trunc | ios::binary);
            ok = ok && cov[i].is_open() && fre[i].is_open();
        }
        return ok;
    }
    void append(pfdropin::CallerFiles &f) {
        allfre << f.allele_frequency; alignseq << f.alignseq;
        f.allele_frequency.clear(); f.alignseq.clear();
        for (int i = 0; i < 4; i++) { cov[i] << f.cov[i]; fre[i] << f.fre[i]; f.cov[i].clear(); f.fre[i].clear(); }
        f.called.clear();
    }
};

bool g_thread_dialect = false;
unsigned g_host_threads = 1;

}  // namespace

void CCDBG::ploidyEstimation_multithread_ptr(const string &outpre, const vector<pair<int, int>> &cutoff, const size_t &thr) {
    g_thread_dialect = thr > 1;
    g_host_threads = (unsigned)max<size_t>(thr, 1);
    ploidyEstimation_ptr(outpre, cutoff);
}

void CCDBG::ploidyEstimation_ptr(const string &outpre, const vector<pair<int, int>> &cutoff) {
    const bool thread_dialect = g_thread_dialect;
    const unsigned T = max(1u, g_host_threads);
    const clock_t start_clock = clock();
    const double start_time = time(NULL);
    const auto t_begin = chrono::steady_clock::now();
    cout << "CCDBG::PloidyEstimation():  Analyzing superbubbles to generate sites' information" << endl;
    if (access("PloidyFrost_output", 0)) { if (system("mkdir ./PloidyFrost_output")) {} }
    g_colored_device.ready();
    const size_t n_colors = cdbg.getNbColors();
    if (!g_colored_device.error.empty() || !g_colored_device.ctx) { cout << "CCDBG::PloidyEstimation(): " << g_colored_device.error << endl; exit(EXIT_FAILURE); }
    if (g_colored_device.dbs.size() < n_colors || cutoff.size() < n_colors) {
        cout << "CCDBG::PloidyEstimation(): " << n_colors << " colours, " << g_colored_device.dbs.size() << " kmc databases, " << cutoff.size() << " thresholds" << endl;
        exit(EXIT_FAILURE);
    }
    Output files_out;
    if (!files_out.open_all("PloidyFrost_output/" + outpre)) { cout << "CCDBG:: PloidyEstimation():Open file error" << endl; exit(EXIT_FAILURE); }
    const size_t k = (size_t)cdbg.getK();
    vector<pf_kmc *> dbs(g_colored_device.dbs.begin(), g_colored_device.dbs.begin() + n_colors);
    vector<pair<int, int>> gates(cutoff.begin(), cutoff.begin() + n_colors);
    ColoredCDBG<MyUnitig> &graph = cdbg;
    pfdropin::ColoredBubbleCaller caller(g_colored_device.ctx, dbs, match, mismatch, gap, gates, (unsigned)k,
        [&graph, k, n_colors](const char *kmer) -> uint64_t {                                  // CCDBG.cpp:1127-1129
            const UCM pu = graph.findUnitig(kmer, 0, k);
            if (pu.isEmpty) return 0;
            const UnitigColors *uc = pu.getData()->getUnitigColors(pu);
            uint64_t mask = 0;
            for (size_t c = 0; c < n_colors; c++) if (uc->contains(pu, c)) mask |= 1ull << c;
            return mask;
        });
    caller.set_thread_dialect(thread_dialect);
    caller.set_host_threads(T);
    {   // bubbles beyond the device limits: the reference's own SeqAlign (CCDBG.cpp:741-753), everything else still batched
        double m_ = match, d_ = mismatch, g_ = gap;
        caller.set_host_aligner([m_, d_, g_](std::vector<std::string> &str, pfdropin::HostMsa &out) mutable {
            SeqAlign seqalign(m_, d_, g_);
            seqalign.SequenceAlignment(str, out.snp_pos, out.indel_pos, out.partition, out.indel_len);
            out.rows = str;
        });
    }
    pfdropin::CallerFiles files;
    size_t var_id = thread_dialect ? 0 : 1;
    size_t n_bubbles = 0;
    double t_collect = 0, t_device_wait = 0, t_call = 0;

    auto device_stage = [&](pfdropin::ColoredBatch &blk) {
        const auto t_s = chrono::steady_clock::now();
        if (!caller.call(blk, var_id, files)) { cout << "CCDBG::PloidyEstimation(): " << caller.error() << endl; exit(EXIT_FAILURE); }
        n_bubbles += blk.flat.n_bubbles();
        files_out.append(files);
        t_call += seconds_since(t_s);
    };

    const size_t kBlock = 1u << 18;
    vector<UCM> units;
    units.reserve(kBlock);
    vector<ThreadOut> outs(T);
    pfdropin::ColoredBatch blocks[2];
    int cur = 0;
    future<void> pending;
    size_t nb_unitig_processed = 0;
    auto it = cdbg.begin();
    const auto it_end = cdbg.end();
    while (it != it_end) {
        auto t_phase = chrono::steady_clock::now();
        units.clear();
        for (; it != it_end && units.size() < kBlock; ++it) {
            units.emplace_back(*it);
            if (++nb_unitig_processed % 100000 == 0) cout << "CCDBG::PloidyEstimation(): Processed " << nb_unitig_processed << " unitigs " << endl;
        }
        const size_t n = units.size();
        const unsigned Tn = (unsigned)min<size_t>(T, max<size_t>(1, n / 256));
        if (Tn == 1) speculate(units, 0, n, k, n_colors, outs[0]);
        else {
            vector<thread> th;
            for (unsigned t = 0; t < Tn; t++) th.emplace_back([&, t] { speculate(units, n * t / Tn, n * (t + 1) / Tn, k, n_colors, outs[t]); });
            for (thread &w : th) w.join();
        }
        for (unsigned t = Tn; t < T; t++) outs[t].clear();
        resolve(outs, blocks[cur]);
        t_collect += seconds_since(t_phase);
        t_phase = chrono::steady_clock::now();
        if (pending.valid()) pending.get();
        t_device_wait += seconds_since(t_phase);
        pfdropin::ColoredBatch *blk = &blocks[cur];
        pending = async(launch::async, [&device_stage, blk] { device_stage(*blk); });
        cur ^= 1;
    }
    {
        const auto t_phase = chrono::steady_clock::now();
        if (pending.valid()) pending.get();
        t_device_wait += seconds_since(t_phase);
    }
    g_colored_device.release();

    const time_t end_time = time(NULL);
    cout << "CCDBG::PloidyEstimation(): Cpu time : " << (double)(clock() - start_clock) / CLOCKS_PER_SEC << "s" << endl;
    cout << "CCDBG::PloidyEstimation(): Real time : " << (double)difftime(end_time, start_time) << "s" << endl;
    {
        const pfdropin::CallerStats &cs = caller.stats();
        cout << "CCDBG::PloidyEstimation():  GPU path : " << n_bubbles << " bubbles, " << n_colors << " colours, " << T << " host threads, phase " << seconds_since(t_begin)
             << "s = collecting " << t_collect << "s, waiting for the device " << t_device_wait << "s; device thread " << t_call << "s (readCovUni " << cs.lookup_s
             << ", gate + order " << cs.gate_s << ", pf_align " << cs.align_s << ", site k-mers + colours + readCov " << cs.site_s << ", rows " << cs.emit_s << "; "
             << cs.bubbles_aligned << " of " << cs.bubbles_in << " bubbles aligned in " << cs.calls << " batches, " << cs.bubbles_host_aligned << " by the host aligner)" << endl;
    }
    cout << "CCDBG::PloidyEstimation(): Alleles in SuperBubbles  :\t"
         << "2 :" << files.alleles[0] << "\t" << "3 :" << files.alleles[1] << "\t" << "4 :" << files.alleles[2] << "\t" << "5 :" << files.alleles[3] << endl;
    if (caller.core_num() != 0) {
        const int avg = (int)(caller.core_cov() / caller.core_num());
        cout << "CCDBG::PloidyEstimation(): Sites' Average Coverage:" << avg << endl;
    }
}
