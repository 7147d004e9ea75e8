// integration/ploidy_estimation_gpu.cpp -- the reference-side binding, as a maintainer of PloidyFrost would add it.
//
// Replacement DEFINITIONS of CDBG::ploidyEstimation_ptr (`-t 1`, declared in the reference's src/CDBG.hpp:39, defined in
// src/CDBG.cpp:1101) and CDBG::ploidyEstimation_multithread_ptr (`-t N`, CDBG.cpp:1872): the per-superbubble analysis of the
// reference -- readCov of the entrance and of the branches, SeqAlign of the branches, site k-mers, output rows -- handed to
// libpfgpu.so in flat batches through pfdropin::BubbleCaller (include/pf_caller.hpp).  Nothing else of the reference changes:
// integration/Makefile compiles the reference's own sources where they lie, weakens the two symbols in its CDBG object and links
// this file + libpfgpu.so into `PloidyFrost_gpu`.  tests/test_gpu_integration.py runs `PloidyFrost` and `PloidyFrost_gpu` on the
// same graph and database and compares every output file byte for byte.
//
// How the phase is organised (the reference does everything per bubble, inside the walk; here the walk only COLLECTS):
//
//   block of unitigs ─► speculate (T threads) ─► resolve (in graph order) ─► flat batch ─► device + rows (own thread) ─► files
//                       └───────────────── block b + 1 ──────────────────┘             └──────────── block b ───────────┘
//
//   * speculate: for every (unitig, strand) that still carries a superbubble pointer, everything that does not depend on other
//     bubbles: is it complex, where does the bubble end, which end reports it (the larger unitig string, CDBG.cpp:1190 / :1349),
//     and the branch strings -- written straight into a per-thread flat arena (include/pf_caller.hpp: FlatBatch), not into
//     std::string vectors.  Contiguous ranges of the block, one host thread each.
//   * resolve: the reference's visited marks (MyUnitig plus/minus bits, CDBG.cpp:1143-1200, :1656-1680) are the only coupling
//     between bubbles: reporting a bubble marks the far strand of its exit, which suppresses whatever would have been opened
//     from there.  They are applied in the reference's iteration order over the speculated records -- a few nanoseconds per
//     record, no graph queries -- so the bubbles, their order and their ids are exactly those of the `-t 1` walk.
//   * device + rows: lookups of the entrances (the reference's readCov(u), including its exit on a missing k-mer), then
//     BubbleCaller::call; runs on its own thread while the host threads speculate the next block.
#include <unistd.h>

#include <chrono>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <future>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "CDBG.hpp"        // the reference's class (Bifrost graph, MyUnitig marks)
#include "SeqAlign.hpp"    // the reference's own aligner: takes the bubbles that exceed a device limit
#include "pf_caller.hpp"   // ours

using namespace std;

namespace {

// The reference opens the KMC database in CDBG's constructor and does not keep its name; a maintainer would store it in the
// class.  From outside the class the name is taken from the command line the way Main.cpp's getopt loop reads it: `-d <prefix>`
// or the attached form `-d<prefix>`, the last occurrence wins.
vector<string> args_of_this_process() {
    ifstream f("/proc/self/cmdline", ios::binary);
    vector<string> args;
    string cur;
    char c;
    while (f.get(c)) { if (c == '\0') { args.push_back(cur); cur.clear(); } else cur += c; }
    return args;
}
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
string kmc_prefix_of_this_process() { return option_of_this_process('d'); }

// Creating the CUDA context takes 0.6 - 2 s and staging the database scales with its size; the reference spends at least as
// long loading the graph and finding the superbubbles before the estimation phase starts.  So both are started on a second
// thread when the program starts (static initialiser, only for command lines that name a graph and a database) and the
// estimation phase joins it.
struct DeviceWarmup {
    thread worker;
    pf_ctx *ctx = nullptr;
    pf_kmc *db = nullptr;
    string error;

    void open_now(const string &prefix) {
        if (pf_init(0, &ctx) != PF_OK) { error = pf_last_error(); ctx = nullptr; return; }
        if (pf_kmc_open(ctx, prefix.c_str(), &db) != PF_OK) { error = pf_last_error(); db = nullptr; return; }
        if (!getenv("PF_NO_WARM")) warm_kernels();
    }
    // CUDA loads a kernel's code at its first launch and the library creates its streams and attributes at first use: a dummy batch --
    // one two-branch bubble per size class plus a few thousand SNP-sized ones -- goes through the three calls here, on the warm-up
    // thread, so that none of that lands in the estimation phase (0.43 -> 0.35 s on the 20 Mbp diploid, profiles/r02_summary.md section 8;
    // a block-sized dummy batch bought nothing more, and in the coloured binding -- eight databases to open first -- it was still
    // running when the phase began).  Results are discarded (the made-up k-mers are simply not in the database).
    void warm_kernels() {
        static const int lens[] = {40, 90, 120, 180, 250, 300};
        const int n_snp = 4096;
        string bases;
        vector<uint64_t> off{0};
        vector<uint32_t> boff{0};
        unsigned x = 12345;
        bases.reserve((size_t)n_snp * 100 + 4096);
        auto add_bubble = [&](int L) {
            const size_t a0 = bases.size();
            for (int i = 0; i < L; i++) { x = x * 1664525u + 1013904223u; bases += "ACGT"[(x >> 24) & 3]; }
            off.push_back(bases.size());
            bases.append(bases, a0, (size_t)L);
            bases[a0 + L + L / 2] = bases[a0 + L / 2] == 'A' ? 'C' : 'A';
            off.push_back(bases.size());
            boff.push_back((uint32_t)(off.size() - 1));
        };
        for (int L : lens) add_bubble(L);
        for (int i = 0; i < n_snp; i++) add_bubble(49);
        vector<pf_cov_t> cov(off.size() - 1);
        pf_msa_batch_t m;
        pf_site_batch_t sc;
        pf_kmc_cov(db, bases.data(), off.data(), (uint32_t)cov.size(), PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu, cov.data());
        if (pf_align(ctx, 2.0, -1.0, -3.0, bases.data(), off.data(), boff.data(), (uint32_t)(boff.size() - 1), &m) == PF_OK)
            pf_site_cov(db, 0, 0xFFFFFFFFu, nullptr, &sc);
    }
    DeviceWarmup() {
        const string prefix = kmc_prefix_of_this_process();
        // coloured command lines (`-f`) name a LIST of databases with `-d`; ploidy_estimation_colored_gpu.cpp brings those up
        if (!prefix.empty() && !option_of_this_process('g').empty() && option_of_this_process('f').empty())
            worker = thread([this, prefix] { open_now(prefix); });
    }
    void ready(const string &prefix) {               // called by the estimation phase
        if (worker.joinable()) worker.join();
        else if (!ctx && error.empty()) open_now(prefix);
    }
    void release() {
        if (db) pf_kmc_close(db);
        if (ctx) pf_shutdown(ctx);
        db = nullptr; ctx = nullptr;
    }
    ~DeviceWarmup() {
        if (worker.joinable()) worker.join();
        release();
    }
};
DeviceWarmup g_device;

double seconds_since(const chrono::steady_clock::time_point &t0) {
    return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}

// ---- speculation: one record per (unitig, strand) that may open a bubble ---------------------------------------------------
enum SpecKind : uint8_t { SPEC_COMPLEX = 0, SPEC_OTHER_END = 1, SPEC_BUBBLE = 2 };
struct Spec {
    MyUnitig *ud, *exit_ud;
    uint8_t strand, exit_strand, kind;
    uint32_t bubble;      // SPEC_BUBBLE: index in the thread's FlatBatch
    uint32_t entrance;    // not SPEC_COMPLEX: index in the thread's entrance arena
};
struct ThreadOut {
    vector<Spec> specs;
    pfdropin::FlatBatch flat;
    string ent_bases;                 // referenceUnitigToString() of every evaluated opener (readCov(u), CDBG.cpp:66-120)
    vector<uint64_t> ent_off{0};
    void clear() { specs.clear(); flat.clear(); ent_bases.clear(); ent_off.assign(1, 0); }
};

// every path from the entrance to the exit, spelled from the entrance's last k-mer to the exit's first k-mer (CDBG.cpp:2226):
// depth first, a unitig's successors taken last-to-first (the order in which the reference's stack hands them out)
void spell_paths(const UnitigMap<MyUnitig> &cur, const UnitigMap<MyUnitig> &exit_uni, size_t k, string &text, pfdropin::FlatBatch &out) {
    vector<UnitigMap<MyUnitig>> next;
    for (const auto &nx : cur.getSuccessors()) next.push_back(nx);
    for (size_t i = next.size(); i-- > 0;) {
        const UnitigMap<MyUnitig> &v = next[i];
        const string s = v.mappedSequenceToString();
        const size_t mark = text.size();
        if (v.isSameReferenceUnitig(exit_uni)) {
            text.push_back(s[k - 1]);                       // the exit contributes the last base of its first k-mer
            out.add_branch(text.data(), text.size(), true);
        } else {
            text.append(s, k - 1, string::npos);            // everything after the (k-1)-base overlap
            spell_paths(v, exit_uni, k, text, out);
        }
        text.resize(mark);
    }
}

void speculate(vector<UnitigMap<MyUnitig>> &units, size_t i0, size_t i1, size_t k, ThreadOut &o) {
    o.clear();
    string text;
    for (size_t i = i0; i < i1; i++) {
        UnitigMap<MyUnitig> u = units[i];
        MyUnitig *ud = u.getData();
        for (int pass = 0; pass < 2; pass++) {
            const bool strand = pass == 0;
            if (strand ? ud->is_plus_visited() : ud->is_minus_visited()) continue;   // no superbubble opens here (or already dealt with)
            Spec sp;
            sp.ud = ud; sp.exit_ud = nullptr; sp.strand = strand; sp.exit_strand = 0; sp.bubble = 0; sp.entrance = 0;
            if (ud->isComplex(strand)) { sp.kind = SPEC_COMPLEX; o.specs.push_back(sp); continue; }
            u.strand = strand;
            const bool strict = ud->isStrict(strand);
            UnitigMap<MyUnitig> exit_uni;
            if (strict) exit_uni = *u.getSuccessors().begin()->getSuccessors().begin();
            else {
                const size_t want = ud->get_bubble_id(strand);
                exit_uni = *u.getSuccessors().begin();
                while (exit_uni.getData()->get_id() != want) exit_uni = *exit_uni.getSuccessors().begin();
            }
            sp.exit_ud = exit_uni.getData();
            sp.exit_strand = exit_uni.strand;
            // readCov(u) reads the entrance's forward string, whichever strand opens the bubble (CDBG.cpp:77)
            const string ref = u.referenceUnitigToString();
            sp.entrance = (uint32_t)(o.ent_off.size() - 1);
            o.ent_bases += ref;
            o.ent_off.push_back(o.ent_bases.size());
            // each bubble is reported from one end only: the end whose unitig string is not the smaller one (CDBG.cpp:1190).
            // Two different unitigs differ inside their first k-mer (a k-mer lives in one unitig), so the heads decide.
            bool other_end;
            if (u.isSameReferenceUnitig(exit_uni)) other_end = false;
            else other_end = ref.compare(0, k, exit_uni.getUnitigHead().toString()) < 0;
            if (other_end) { sp.kind = SPEC_OTHER_END; o.specs.push_back(sp); continue; }
            sp.kind = SPEC_BUBBLE;
            sp.bubble = (uint32_t)o.flat.n_bubbles();
            if (strict) {
                for (const auto &uu : u.getSuccessors()) {
                    const string s = uu.mappedSequenceToString();
                    o.flat.add_branch(s.data(), s.size(), uu.strand);
                }
            } else {
                const string s = u.mappedSequenceToString();
                text.assign(s, s.size() - k, k);             // the entrance's last k-mer
                spell_paths(u, exit_uni, k, text, o.flat);
            }
            o.flat.end_bubble(strict, (unsigned)ud->get_id(), (unsigned)sp.exit_ud->get_id(), u.size, exit_uni.size);
            o.specs.push_back(sp);
        }
    }
}

// one block after resolution: what goes to the device
struct Block {
    pfdropin::FlatBatch flat;
    string ent_bases;
    vector<uint64_t> ent_off{0};
    vector<uint32_t> bubble_entrance;   // per bubble of `flat`: its entrance in the arrays above
    void clear() { flat.clear(); ent_bases.clear(); ent_off.assign(1, 0); bubble_entrance.clear(); }
};

void copy_bubble(const pfdropin::FlatBatch &from, size_t b, pfdropin::FlatBatch &to) {
    for (uint32_t s = from.bubble_off[b]; s < from.bubble_off[b + 1]; s++)
        to.add_branch(from.bases.data() + from.seq_off[s], (size_t)(from.seq_off[s + 1] - from.seq_off[s]), from.fwd[s] != 0);
    to.end_bubble(from.strict[b] != 0, from.entrance_id[b], from.exit_id[b], (size_t)from.entrance_size[b], (size_t)from.exit_size[b]);
}

// the reference's marks, in its order (plus strand before minus, unitigs in graph order)
void resolve(vector<ThreadOut> &outs, Block &blk) {
    blk.clear();
    for (ThreadOut &o : outs)
        for (const Spec &sp : o.specs) {
            MyUnitig *ud = sp.ud;
            if (sp.strand ? ud->is_plus_visited() : ud->is_minus_visited()) continue;   // suppressed by a bubble reported before
            if (sp.strand) ud->set_plus_visited(); else ud->set_minus_visited();         // every path of the reference's loop ends here
            if (sp.kind == SPEC_COMPLEX) continue;
            const uint32_t e = (uint32_t)(blk.ent_off.size() - 1);                        // readCov(u) is due (CDBG.cpp:1186)
            blk.ent_bases.append(o.ent_bases, o.ent_off[sp.entrance], o.ent_off[sp.entrance + 1] - o.ent_off[sp.entrance]);
            blk.ent_off.push_back(blk.ent_bases.size());
            if (sp.kind == SPEC_OTHER_END) continue;
            copy_bubble(o.flat, sp.bubble, blk.flat);
            blk.bubble_entrance.push_back(e);
            if (sp.exit_strand) sp.exit_ud->set_minus_visited(); else sp.exit_ud->set_plus_visited();
        }
}

string revcomp(const string &s) {
    string r(s.rbegin(), s.rend());
    for (char &c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    return r;
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
    void append(pfdropin::CallerFiles &f) {          // rows leave as soon as their batch is done
        allfre << f.allele_frequency; alignseq << f.alignseq;
        f.allele_frequency.clear(); f.alignseq.clear();
        for (int i = 0; i < 4; i++) { cov[i] << f.cov[i]; fre[i] << f.fre[i]; f.cov[i].clear(); f.fre[i].clear(); }
        f.called.clear();
    }
};

}  // namespace

// The `-t N` entry (CDBG.hpp:33, CDBG.cpp:1872): N host threads collect and format, the device does the per-bubble work; what
// changes against `-t 1` is the dialect of the files (0-based ids, P_allele_frequency grouped per bubble).  The reference's own
// `-t N` files are schedule-dependent in row order and ids; ours are one legal schedule, always the same.
static bool g_thread_dialect = false;
static unsigned g_host_threads = 1;

void CDBG::ploidyEstimation_multithread_ptr(const string &outpre, const int &lower, const int &upper, const size_t &thr) {
    g_thread_dialect = thr > 1;
    g_host_threads = (unsigned)max<size_t>(thr, 1);
    ploidyEstimation_ptr(outpre, lower, upper);
}

void CDBG::ploidyEstimation_ptr(const string &outpre, const int &lower, const int &upper) {
    const bool thread_dialect = g_thread_dialect;
    const unsigned T = max(1u, g_host_threads);
    const clock_t start_clock = clock();
    const double start_time = time(NULL);
    const auto t_begin = chrono::steady_clock::now();
    cout << "CDBG::PloidyEstimation():  Analyzing superbubbles to generate sites' information" << endl;
    if (access("PloidyFrost_output", 0)) { if (system("mkdir ./PloidyFrost_output")) {} }
    auto t_phase = chrono::steady_clock::now();
    g_device.ready(kmc_prefix_of_this_process());
    if (!g_device.error.empty() || !g_device.db) { cout << "CDBG::PloidyEstimation(): " << g_device.error << endl; exit(EXIT_FAILURE); }
    Output files_out;       // opened before the first bubble is looked at: a directory that cannot be written ends the run at once
    if (!files_out.open_all("PloidyFrost_output/" + outpre)) { cout << "CDBG:: PloidyEstimation():Open file error" << endl; exit(EXIT_FAILURE); }
    pf_ctx *ctx = g_device.ctx;
    pf_kmc *db = g_device.db;
    const double t_open = seconds_since(t_phase);
    pfdropin::BubbleCaller caller(ctx, db, match, mismatch, gap, (unsigned)lower, (unsigned)upper);
    caller.set_thread_dialect(thread_dialect);
    caller.set_host_threads(T);
    // a bubble beyond the device limits (more than 64 co-optimal alignments / candidate MSAs, 64 rows) is aligned by the reference's
    // own SeqAlign, as the reference would have done (CDBG.cpp:2036-2050); everything else about it still goes through the batch
    {
        double m_ = match, d_ = mismatch, g_ = gap;
        caller.set_host_aligner([m_, d_, g_](std::vector<std::string> &str, pfdropin::HostMsa &out) mutable {
            SeqAlign seqalign(m_, d_, g_);
            seqalign.SequenceAlignment(str, out.snp_pos, out.indel_pos, out.partition, out.indel_len);
            out.rows = str;
        });
    }
    pfdropin::CallerFiles files;
    size_t var_id = thread_dialect ? 0 : 1;
    size_t coreNum = 0, coreCov = 0, n_bubbles = 0;
    const size_t k = (size_t)cdbg.getK();
    double t_collect = 0, t_device_wait = 0, t_iter = 0, t_resolve = 0;

    // the device side of one block; runs on its own thread while the next block is collected
    vector<pf_cov_t> ent_cov;
    double t_entrance = 0, t_call = 0, t_write = 0;
    auto device_stage = [&](Block &blk) {
        auto t_s = chrono::steady_clock::now();
        const uint32_t n_ent = (uint32_t)(blk.ent_off.size() - 1);
        ent_cov.resize(n_ent);
        if (n_ent && pf_kmc_cov(db, blk.ent_bases.data(), blk.ent_off.data(), n_ent, PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu, ent_cov.data()) != PF_OK) {
            cout << "CDBG::PloidyEstimation(): " << pf_last_error() << endl;
            exit(EXIT_FAILURE);
        }
        for (uint32_t e = 0; e < n_ent; e++)
            if (ent_cov[e].first_missing >= 0) {        // readCov(u) ends the program on a k-mer the database does not hold (CDBG.cpp:92-96)
                const string km = blk.ent_bases.substr(blk.ent_off[e] + (size_t)ent_cov[e].first_missing, k);
                cout << "CDBG::readCov():" << revcomp(km) << " kmer can not found ." << endl;
                exit(EXIT_FAILURE);
            }
        t_entrance += seconds_since(t_s); t_s = chrono::steady_clock::now();
        if (!caller.call(blk.flat, var_id, files)) { cout << "CDBG::readCov():" << caller.error() << endl; exit(EXIT_FAILURE); }
        t_call += seconds_since(t_s); t_s = chrono::steady_clock::now();
        for (size_t b = 0; b < blk.flat.n_bubbles(); b++)
            if (files.called[b]) {                      // mean coverage of the entrances of the called bubbles (only printed, CDBG.cpp:1261, :1703)
                const pf_cov_t &c = ent_cov[blk.bubble_entrance[b]];
                coreCov += (size_t)((double)c.sum / (double)c.n_kmers);
                coreNum++;
            }
        n_bubbles += blk.flat.n_bubbles();
        files_out.append(files);
        t_write += seconds_since(t_s);
    };

    // ---- blocks of unitigs: collect block b + 1 while the device works on block b ----
    const size_t kBlock = 1u << 18;
    vector<UnitigMap<MyUnitig>> units;
    units.reserve(kBlock);
    vector<ThreadOut> outs(T);
    Block blocks[2];
    int cur = 0;
    future<void> pending;
    size_t nb_unitig_processed = 0;
    auto it = cdbg.begin();
    const auto it_end = cdbg.end();
    while (it != it_end) {
        t_phase = chrono::steady_clock::now();
        units.clear();
        for (; it != it_end && units.size() < kBlock; ++it) {
            units.emplace_back(*it);
            if (++nb_unitig_processed % 100000 == 0) cout << "CDBG::PloidyEstimation(): Processed " << nb_unitig_processed << " unitigs " << endl;
        }
        const size_t n = units.size();
        t_iter += seconds_since(t_phase);
        const unsigned Tn = (unsigned)min<size_t>(T, max<size_t>(1, n / 256));
        if (Tn == 1) speculate(units, 0, n, k, outs[0]);
        else {
            vector<thread> th;
            for (unsigned t = 0; t < Tn; t++) th.emplace_back([&, t] { speculate(units, n * t / Tn, n * (t + 1) / Tn, k, outs[t]); });
            for (thread &w : th) w.join();
        }
        for (unsigned t = Tn; t < T; t++) outs[t].clear();
        const auto t_res = chrono::steady_clock::now();
        resolve(outs, blocks[cur]);
        t_resolve += seconds_since(t_res);
        t_collect += seconds_since(t_phase);
        t_phase = chrono::steady_clock::now();
        if (pending.valid()) pending.get();                 // the device is done with the previous block
        t_device_wait += seconds_since(t_phase);
        Block *blk = &blocks[cur];
        pending = async(launch::async, [&device_stage, blk] { device_stage(*blk); });
        cur ^= 1;
    }
    t_phase = chrono::steady_clock::now();
    if (pending.valid()) pending.get();
    t_device_wait += seconds_since(t_phase);
    t_phase = chrono::steady_clock::now();
    g_device.release();
    const double t_release = seconds_since(t_phase);

    const time_t end_time = time(NULL);
    cout << "CDBG::PloidyEstimation():  Cpu time : " << (double)(clock() - start_clock) / CLOCKS_PER_SEC << "s" << endl;
    cout << "CDBG::PloidyEstimation():  Real time : " << (double)difftime(end_time, start_time) << "s" << endl;
    cout << "CDBG::PloidyEstimation():  GPU path : " << n_bubbles << " bubbles, " << T << " host threads, phase " << seconds_since(t_begin)
         << "s = waited for device + database " << t_open << "s, collecting " << t_collect << "s (unitig iteration " << t_iter << ", resolve " << t_resolve << "), waiting for the device " << t_device_wait << "s" << ", releasing the device " << t_release << "s" << endl;
    {
        const pfdropin::CallerStats &cs = caller.stats();
        cout << "CDBG::PloidyEstimation():  GPU path, device thread : entrance readCov " << t_entrance << "s, BubbleCaller::call " << t_call << "s (branch readCov "
             << cs.lookup_s << ", gate + order " << cs.gate_s << ", pf_align " << cs.align_s << ", pf_site_cov " << cs.site_s << ", rows " << cs.emit_s
             << "; " << cs.bubbles_aligned << " of " << cs.bubbles_in << " bubbles aligned in " << cs.calls << " batches, " << cs.bubbles_host_aligned
             << " by the host aligner), coverage + files " << t_write << "s" << endl;
    }
    cout << "CDBG::PloidyEstimation(): Alleles in SuperBubbles  :\t"
         << "2 :" << files.alleles[0] << "\t" << "3 :" << files.alleles[1] << "\t" << "4 :" << files.alleles[2] << "\t" << "5 :" << files.alleles[3] << endl;
    const int avg = coreNum ? (int)(coreCov / coreNum) : 0;
    cout << "CDBG::PloidyEstimation(): Sites' Average Coverage:" << avg << endl;
}
