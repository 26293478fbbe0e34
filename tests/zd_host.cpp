// Host build of the device zstd decoder (fhe_precompiles_b200/csrc/zstd_dec.h) for tests/test_zstd_dec.py: the same source
// that k_zstd_inflate runs, compiled with g++ so that it can be compared with libzstd and fuzzed without a GPU.
#include <cstring>
#include <vector>

#include "zstd_dec.h"

extern "C" int zd_decode(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    static thread_local fheb::zd::Work *w = nullptr;
    if (!w) {
        w = new fheb::zd::Work();
        fheb::zd::work_bind(w, nullptr);
    }
    std::vector<uint8_t> buf(slen + 2 * fheb::zd::kPad, 0xAA);  // the decoder reads aligned words around the frame
    memcpy(buf.data() + fheb::zd::kPad, src, slen);
    return fheb::zd::decode_frame(buf.data() + fheb::zd::kPad, slen, dst, cap, dlen, w);
}

// the two-phase pipeline (plan_frame + execute_plan), scalar on the host
extern "C" int zd_decode_two_phase(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    static thread_local fheb::zd::Work *w = nullptr;
    static thread_local fheb::zd::FramePlan *plan = nullptr;
    static thread_local std::vector<uint64_t> seqs(fheb::zd::kPlanMaxSeqs);
    if (!w) {
        w = new fheb::zd::Work();
        fheb::zd::work_bind(w, nullptr);
        plan = new fheb::zd::FramePlan();
    }
    std::vector<uint8_t> buf(slen + 2 * fheb::zd::kPad, 0xAA), lits(cap + 8);
    memcpy(buf.data() + fheb::zd::kPad, src, slen);
    const uint8_t *f = buf.data() + fheb::zd::kPad;
    if (fheb::zd::plan_frame(f, slen, cap, w, plan, seqs.data(), lits.data()) != fheb::zd::kZdOk) return 1;
    return fheb::zd::execute_plan(f, plan, seqs.data(), lits.data(), dst, dlen);
}

// the batch-oriented pipeline (zstd_plan2.h): parse -> four Huffman stream "threads" + the sequence "thread" -> execution
#include "zstd_plan2.h"
extern "C" int zd_decode_v2(const uint8_t *src, size_t slen, uint8_t *dst, size_t cap, size_t *dlen) {
    using namespace fheb::zd;
    static thread_local Work *w = nullptr;
    static thread_local Plan2 *plan = nullptr;
    static thread_local Tables2 *tabs = nullptr;
    static thread_local std::vector<uint64_t> seqs(kP2MaxSeqs);
    if (!w) {
        w = new Work();
        plan = new Plan2();
        tabs = new Tables2();
    }
    work_bind(w, nullptr);
    // an odd leading pad so that stream starts land on every alignment
    static thread_local unsigned shift = 0;
    shift = (shift + 1) & 7;
    std::vector<uint8_t> buf(slen + 2 * kPad + 8, 0xAA), lits(cap + 16);
    uint8_t *f = buf.data() + kPad + shift;
    memcpy(f, src, slen);
    if (plan2_parse(f, slen, cap, w, plan, tabs) != kZdOk) return 1;
    for (int k = 0; k < 4; k++) plan->huf_bad[k] = plan2_huf(f, plan, tabs, lits.data(), k) ? 0 : 1;
    plan->seq_bad = plan2_seq(f, plan, tabs, seqs.data()) ? 0 : 1;
    // the executor stores aligned 64-bit words: an aligned bounce buffer with room up to the next multiple of 8
    std::vector<uint64_t> out((cap + 15) / 8);
    const int rc = plan2_exec(f, plan, seqs.data(), lits.data(), (uint8_t *)out.data(), dlen);
    if (rc == 0) memcpy(dst, out.data(), *dlen);
    return rc;
}
