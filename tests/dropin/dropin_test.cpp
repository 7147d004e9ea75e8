// tests/dropin/dropin_test.cpp -- TEST: the C++ look-alike adapters of include/pf_dropin.hpp, used the way the
// reference uses CKMCFile / SeqAlign (CDBG.cpp:29-57, :2036-2050).  Built by tests/test_gpu_dropin.py with g++
// against libpfgpu.so; prints one line per check, exit code 0 = all passed.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pf_dropin.hpp"

static int fails = 0;
#define CHECK(cond, what) do { if (!(cond)) { printf("FAIL %s\n", what); fails++; } else printf("ok   %s\n", what); } while (0)

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: dropin_test <kmc prefix> <present 25-mer> <absent 25-mer>\n"); return 2; }
    // --- CKMCFile-style use ---
    pfdropin::KmcFile db;
    CHECK(db.OpenForRA(argv[1]), "OpenForRA");
    CHECK(!db.OpenForRA(argv[1]), "second OpenForRA refused");
    pfdropin::KmcFileInfo info;
    CHECK(db.Info(info) && info.kmer_length == 25 && info.counter_size == 2, "Info");
    CHECK(db.KmerLength() == 25 && db.GetBothStrands(), "KmerLength / GetBothStrands");
    std::string present = argv[2], absent = argv[3];
    unsigned int cnt = 0;
    // the reference's readCov pattern: if (!IsKmer(k)) k.reverse(); CheckKmer(k, cnt)   (CDBG.cpp:38-43)
    std::string rc(present.rbegin(), present.rend());
    for (char &c : rc) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
    const bool fwd = db.IsKmer(present), rev = db.IsKmer(rc);
    CHECK(fwd || rev, "IsKmer on one strand");
    CHECK(db.CheckKmer(fwd ? present : rc, cnt) && cnt >= 1, "CheckKmer count");
    CHECK(!db.IsKmer(absent), "absent k-mer");
    std::vector<uint32_t> counters;
    CHECK(db.GetCountersForRead(present + "ACGT", counters) && counters.size() == 5 && counters[0] == cnt, "GetCountersForRead");
    CHECK(!db.GetCountersForRead("ACGT", counters) && counters.empty(), "GetCountersForRead short read");
    std::vector<pf_cov_t> cov;
    CHECK(db.ReadCovBatch({present, absent, "ACGT"}, 0, 100000, cov) && cov[0].sum == cnt && cov[0].first_missing == -1 &&
              cov[1].first_missing == 0 && cov[2].n_kmers == 0, "ReadCovBatch");
    CHECK(db.Close() && !db.Close(), "Close");
    // --- SeqAlign-style use (literal vectors of SURVEY.md section 4) ---
    double M = 2, D = -1, G = -3;
    pfdropin::SeqAlign sa(M, D, G);
    std::vector<std::string> str = {"ACGTAAAATTGCA", "ACGTAAATTGCA", "ACGTCAAATTGCA"};
    std::vector<unsigned int> snp_pos, indel_pos, indel_len;
    std::vector<std::vector<unsigned short>> partition;
    sa.SequenceAlignment(str, snp_pos, indel_pos, partition, indel_len);
    CHECK(str.size() == 3 && str[1] == "ACGTAAA-TTGCA", "SequenceAlignment rows");
    CHECK(snp_pos == std::vector<unsigned int>{4} && indel_pos == std::vector<unsigned int>{7} && indel_len == std::vector<unsigned int>{1}, "site lists");
    CHECK(partition.size() == 13 && partition[4] == (std::vector<unsigned short>{1, 1, 2}) && partition[7] == (std::vector<unsigned short>{1, 2, 1}) &&
              partition[0].back() == 0, "partition");
    std::vector<std::vector<std::string>> bubbles = {{"ACGTACGTAC", "ACGTTCGTAC"}, {"AAAAAAAAAA", "AAAAAAAA"}, {"ACGT"}};
    std::vector<pfdropin::MsaOut> out;
    std::vector<int> status;
    CHECK(sa.SequenceAlignmentBatch(bubbles, out, status) && out[0].snp_pos == std::vector<unsigned int>{4} && bubbles[1][1] == "AAAAAAAA--" &&
              out[1].indel_len.empty() && status[2] == PF_BUBBLE_BAD_INPUT && bubbles[2].empty(), "SequenceAlignmentBatch");
    printf("%s\n", fails ? "FAILED" : "ALL OK");
    return fails ? 1 : 0;
}
