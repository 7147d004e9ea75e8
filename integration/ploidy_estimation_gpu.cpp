// integration/ploidy_estimation_gpu.cpp -- the reference-side binding, as a maintainer of PloidyFrost would add it.
//
// A replacement DEFINITION of CDBG::ploidyEstimation_ptr (the `-t 1` entry of the per-superbubble analysis, declared in the
// reference's src/CDBG.hpp:39 and defined in src/CDBG.cpp:1101): it keeps the reference's walk over its Bifrost graph -- which
// unitig / strand opens a bubble, where the bubble ends, the orientation rule, the visited marks (CDBG.cpp:1143-1200,
// :1345-1410, :1656-1680) -- but instead of calling readCov / SeqAlign per bubble it collects pfdropin::Bubble records and hands
// them to pfdropin::BubbleCaller (include/pf_caller.hpp -> libpfgpu.so) in batches; the text that comes back is written to the
// same files.  Nothing else of the reference changes: integration/Makefile compiles the reference's own sources where they lie,
// weakens the one symbol in its CDBG object and links this file + libpfgpu.so into `PloidyFrost_gpu`.
// tests/test_gpu_integration.py runs `PloidyFrost` and `PloidyFrost_gpu` on the same graph and database and compares every
// output file byte for byte.
#include <unistd.h>

#include <chrono>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <stack>
#include <string>
#include <thread>
#include <vector>

#include "CDBG.hpp"        // the reference's class (Bifrost graph, MyUnitig marks)
#include "pf_caller.hpp"   // ours

using namespace std;

namespace {

// The reference opens the KMC database in CDBG's constructor and does not keep its name; a maintainer would store it in the
// class.  From outside the class the name is taken from the command line (`-d <prefix>`).
string option_of_this_process(const string &flag) {
    ifstream f("/proc/self/cmdline", ios::binary);
    vector<string> args;
    string cur;
    char c;
    while (f.get(c)) { if (c == '\0') { args.push_back(cur); cur.clear(); } else cur += c; }
    for (size_t i = 0; i + 1 < args.size(); i++)
        if (args[i] == flag) return args[i + 1];
    return "";
}
string kmc_prefix_of_this_process() { return option_of_this_process("-d"); }

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
        if (pf_kmc_open(ctx, prefix.c_str(), &db) != PF_OK) { error = pf_last_error(); db = nullptr; }
    }
    DeviceWarmup() {
        const string prefix = kmc_prefix_of_this_process();
        if (!prefix.empty() && !option_of_this_process("-g").empty()) worker = thread([this, prefix] { open_now(prefix); });
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

void write_text(const string &path, const string &text) {
    ofstream f(path, ios::out | ios::trunc | ios::binary);
    if (!f.is_open()) { cout << "CDBG:: PloidyEstimation():Open file error" << endl; exit(EXIT_FAILURE); }
    f << text;
}

}  // namespace

// The `-t N` entry (CDBG.hpp:33, CDBG.cpp:1872): the device takes the place of the worker threads, so the walk is the same
// single pass; what changes is the dialect of the files (0-based ids, P_allele_frequency grouped per bubble).  The reference's
// own `-t N` files are schedule-dependent in row order and ids; ours are one legal schedule, always the same.
static bool g_thread_dialect = false;
static unsigned g_host_threads = 1;   // -t N: the threads format the rows of a batch (BubbleCaller::set_host_threads)

void CDBG::ploidyEstimation_multithread_ptr(const string &outpre, const int &lower, const int &upper, const size_t &thr) {
    g_thread_dialect = thr > 1;
    g_host_threads = (unsigned)thr;
    ploidyEstimation_ptr(outpre, lower, upper);
}

void CDBG::ploidyEstimation_ptr(const string &outpre, const int &lower, const int &upper) {
    const bool thread_dialect = g_thread_dialect;
    const clock_t start_clock = clock();
    const double start_time = time(NULL);
    cout << "CDBG::PloidyEstimation():  Analyzing superbubbles to generate sites' information" << endl;
    if (access("PloidyFrost_output", 0)) { if (system("mkdir ./PloidyFrost_output")) {} }

    // ---- the walk: which bubbles, in which order, with which marks ----
    auto t_phase = chrono::steady_clock::now();
    vector<pfdropin::Bubble> bubbles;
    vector<string> entrance_seq;
    size_t nb_unitig_processed = 0;
    for (const auto &unitig : cdbg) {
        ++nb_unitig_processed;
        if (nb_unitig_processed % 100000 == 0) cout << "CDBG::PloidyEstimation(): Processed " << nb_unitig_processed << " unitigs " << endl;
        UnitigMap<MyUnitig> u(unitig);
        MyUnitig *ud = u.getData();
        if (ud->is_both_visited()) continue;
        while (!ud->is_both_visited()) {
            if (!ud->is_plus_visited()) {
                u.strand = true;
                if (ud->isComplex(u.strand)) { ud->set_plus_visited(); continue; }
            } else if (!ud->is_minus_visited()) {
                u.strand = false;
                if (ud->isComplex(u.strand)) { ud->set_minus_visited(); break; }
            } else break;
            const bool strict = ud->isStrict(u.strand);
            UnitigMap<MyUnitig> exit_uni;
            if (strict) exit_uni = (*u.getSuccessors().begin()->getSuccessors().begin());
            else {
                exit_uni = *u.getSuccessors().begin();
                while (exit_uni.getData()->get_id() != ud->get_bubble_id(u.strand)) exit_uni = *exit_uni.getSuccessors().begin();
            }
            if (u.referenceUnitigToString().compare(exit_uni.referenceUnitigToString()) < 0) {   // each bubble is reported from one side only
                if (u.strand) ud->set_plus_visited(); else ud->set_minus_visited();
                continue;
            }
            pfdropin::Bubble b;
            b.strict = strict;
            b.entrance_id = (unsigned)ud->get_id();
            b.exit_id = (unsigned)exit_uni.getData()->get_id();
            b.entrance_size = u.size;
            b.exit_size = exit_uni.size;
            if (strict) {
                for (const auto &uu : u.getSuccessors()) {
                    b.branches.push_back(uu.mappedSequenceToString());
                    b.sort_keys.push_back(uu.referenceUnitigToString());
                }
            } else {
                // every path from the entrance to the exit, as the string from the entrance's last k-mer to the exit's first
                // k-mer: depth-first over the successors with the unitigs of the current path on `path` and its text in `text`
                stack<UnitigMap<MyUnitig>> path, todo;
                string text;
                todo.push(u);
                while (!todo.empty()) {
                    UnitigMap<MyUnitig> cur = todo.top();
                    todo.pop();
                    path.push(cur);
                    const string str = cur.mappedSequenceToString();
                    text += str.substr(0, cur.len);
                    if (cur.isSameReferenceUnitig(exit_uni)) {
                        text += str.substr(cur.len);
                        b.branches.push_back(text.substr(u.len - 1, text.length() - u.len + 1 - cur.len + 1));
                        text = text.substr(0, text.length() - str.length());
                        path.pop();
                        while (!path.empty() && !todo.empty()) {      // unwind to the unitig the next pending one hangs off
                            bool parent = false;
                            for (const auto &nx : path.top().getSuccessors())
                                if (nx == todo.top()) { parent = true; break; }
                            if (parent) break;
                            text = text.substr(0, text.length() - path.top().len);
                            path.pop();
                        }
                    } else {
                        for (const auto &nx : cur.getSuccessors()) todo.push(nx);
                    }
                }
            }
            bubbles.push_back(std::move(b));
            entrance_seq.push_back(u.referenceUnitigToString());
            if (u.strand) ud->set_plus_visited(); else ud->set_minus_visited();
            if (exit_uni.strand) exit_uni.getData()->set_minus_visited(); else exit_uni.getData()->set_plus_visited();
        }
    }

    // ---- the device: lookups, alignment, site k-mers for all bubbles, in batches ----
    const double t_walk = seconds_since(t_phase);
    t_phase = chrono::steady_clock::now();
    g_device.ready(kmc_prefix_of_this_process());
    if (!g_device.error.empty() || !g_device.db) { cout << "CDBG::PloidyEstimation(): " << g_device.error << endl; exit(EXIT_FAILURE); }
    pf_ctx *ctx = g_device.ctx;
    pf_kmc *db = g_device.db;
    const double t_open = seconds_since(t_phase);
    t_phase = chrono::steady_clock::now();
    pfdropin::BubbleCaller caller(ctx, db, match, mismatch, gap, (unsigned)lower, (unsigned)upper);
    caller.set_thread_dialect(thread_dialect);
    caller.set_host_threads(g_host_threads);
    pfdropin::CallerFiles files;
    size_t var_id = thread_dialect ? 0 : 1;
    const size_t kBatch = 1u << 18;
    for (size_t at = 0; at < bubbles.size(); at += kBatch) {
        if (!caller.call(bubbles.data() + at, min(bubbles.size() - at, kBatch), var_id, files)) { cout << "CDBG::readCov():" << caller.error() << endl; exit(EXIT_FAILURE); }
    }
    // mean coverage of the entrances of the called bubbles (only printed, CDBG.cpp:1186, :1703)
    size_t coreNum = 0, coreCov = 0;
    {
        string flat;
        vector<uint64_t> off(1, 0);
        for (size_t i = 0; i < bubbles.size(); i++)
            if (files.called[i]) { flat += entrance_seq[i]; off.push_back(flat.size()); }
        vector<pf_cov_t> cov(off.size() - 1);
        if (!cov.empty()) {
            if (pf_kmc_cov(db, flat.data(), off.data(), (uint32_t)cov.size(), PF_LOOKUP_FWD_THEN_RC, 0, 0xFFFFFFFFu, cov.data()) != PF_OK) {
                cout << "CDBG::PloidyEstimation(): " << pf_last_error() << endl;
                exit(EXIT_FAILURE);
            }
            for (const pf_cov_t &c : cov) { coreCov += (size_t)((double)c.sum / (double)c.n_kmers); coreNum++; }
        }
    }
    const double t_calls = seconds_since(t_phase);
    g_device.release();

    // ---- the files ----
    const string dir = "PloidyFrost_output/" + outpre;
    static const char *names[4] = {"bi", "tri", "tetra", "penta"};
    write_text(dir + "_allele_frequency.txt", files.allele_frequency);
    write_text(dir + "_alignseq.txt", files.alignseq);
    for (int i = 0; i < 4; i++) {
        write_text(dir + "_" + names[i] + "cov.txt", files.cov[i]);
        write_text(dir + "_" + names[i] + "fre.txt", files.fre[i]);
    }
    const time_t end_time = time(NULL);
    cout << "CDBG::PloidyEstimation():  Cpu time : " << (double)(clock() - start_clock) / CLOCKS_PER_SEC << "s" << endl;
    cout << "CDBG::PloidyEstimation():  Real time : " << (double)difftime(end_time, start_time) << "s" << endl;
    cout << "CDBG::PloidyEstimation():  GPU path : " << bubbles.size() << " bubbles, graph walk " << t_walk << "s, waited for device + database "
         << t_open << "s, lookups + alignment + rows " << t_calls << "s" << endl;
    cout << "CDBG::PloidyEstimation(): Alleles in SuperBubbles  :\t"
         << "2 :" << files.alleles[0] << "\t" << "3 :" << files.alleles[1] << "\t" << "4 :" << files.alleles[2] << "\t" << "5 :" << files.alleles[3] << endl;
    const int avg = coreNum ? (int)(coreCov / coreNum) : 0;
    cout << "CDBG::PloidyEstimation(): Sites' Average Coverage:" << avg << endl;
}
