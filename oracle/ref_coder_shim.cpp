// TEST INFRASTRUCTURE ONLY (oracle). Thin C shim over the UNMODIFIED reference host coder classes
// (ArithmeticEncoder / ArithmeticDecoder / BitOutputStream / BitInputStream), which are compiled from
// /root/reference/extension/{ArithmeticCoder,BitIoStream}.cpp in place by oracle/Makefile.ref.
// It restates the four slice loops of the reference `Coder` (coder.cpp:30-47 encodes, :49-69 decodes,
// :70-89 encodes_mask, :90-114 decodes_mask, coder.h:15-35 start/end) over plain pointers, because
// coder.cpp itself takes at::Tensor arguments and cannot be built without torch.
// Streams are in-memory (std::stringstream) instead of files; the byte sequence is what the reference
// writes to its output file.
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include "ArithmeticCoder.h"
#include "BitIoStream.h"

namespace {
struct RefCoder {
    std::stringstream ss;
    BitOutputStream* bout = nullptr;
    ArithmeticEncoder* enc = nullptr;
    BitInputStream* bin = nullptr;
    ArithmeticDecoder* dec = nullptr;
    float fill = 3.5f;
    std::string err;
};
}  // namespace

extern "C" {

void* refcoder_create(float fill) {
    RefCoder* c = new RefCoder();
    c->fill = fill;
    return c;
}

void refcoder_destroy(void* h) {
    RefCoder* c = static_cast<RefCoder*>(h);
    delete c->enc; delete c->bout; delete c->dec; delete c->bin;
    delete c;
}

const char* refcoder_error(void* h) { return static_cast<RefCoder*>(h)->err.c_str(); }

// coder.h:15-21
int refcoder_start_encoder(void* h) {
    RefCoder* c = static_cast<RefCoder*>(h);
    try {
        c->ss.str(std::string()); c->ss.clear();
        delete c->enc; delete c->bout;
        c->bout = new BitOutputStream(c->ss);
        c->enc = new ArithmeticEncoder(32, *c->bout);
    } catch (const char* e) { c->err = e; return 1; }
    return 0;
}

// coder.cpp:30-47 (mask == nullptr) and :70-89 (mask != nullptr)
int refcoder_encode_rows(void* h, const int32_t* table, int ncode, const int32_t* label, const float* mask, int num) {
    RefCoder* c = static_cast<RefCoder*>(h);
    std::vector<uint32_t> row(ncode + 1);
    try {
        for (int i = 0; i < num; i++) {
            if (mask && mask[i] < 0.5f) continue;
            for (int j = 0; j <= ncode; j++) row[j] = static_cast<uint32_t>(table[(size_t)i * (ncode + 1) + j]);
            c->enc->write(row.data(), (uint32_t)ncode, row[ncode], static_cast<uint32_t>(label[i]));
        }
    } catch (const char* e) { c->err = e; return 1; }
    return 0;
}

// coder.h:22-26; returns the number of bytes of the finished stream
long refcoder_end_encoder(void* h) {
    RefCoder* c = static_cast<RefCoder*>(h);
    try {
        c->enc->finish();
        c->bout->finish();
    } catch (const char* e) { c->err = e; return -1; }
    return (long)c->ss.str().size();
}

long refcoder_get_bytes(void* h, uint8_t* out, long cap) {
    RefCoder* c = static_cast<RefCoder*>(h);
    std::string s = c->ss.str();
    long n = (long)s.size() < cap ? (long)s.size() : cap;
    memcpy(out, s.data(), n);
    return (long)s.size();
}

// coder.h:30-35
int refcoder_start_decoder(void* h, const uint8_t* bytes, long n) {
    RefCoder* c = static_cast<RefCoder*>(h);
    try {
        c->ss.str(std::string(reinterpret_cast<const char*>(bytes), (size_t)n)); c->ss.clear();
        delete c->dec; delete c->bin;
        c->bin = new BitInputStream(c->ss);
        c->dec = new ArithmeticDecoder(32, *c->bin);
    } catch (const char* e) { c->err = e; return 1; }
    return 0;
}

// coder.cpp:49-69 (mask == nullptr) and :90-114 (mask != nullptr); out receives `num` floats
int refcoder_decode_rows(void* h, const int32_t* table, int ncode, const float* mask, int num, float* out) {
    RefCoder* c = static_cast<RefCoder*>(h);
    std::vector<uint32_t> row(ncode + 1);
    try {
        for (int i = 0; i < num; i++) {
            if (mask && mask[i] < 0.5f) { out[i] = c->fill; continue; }
            for (int j = 0; j <= ncode; j++) row[j] = static_cast<uint32_t>(table[(size_t)i * (ncode + 1) + j]);
            out[i] = static_cast<float>(c->dec->read(row.data(), (uint32_t)ncode, row[ncode]));
        }
    } catch (const char* e) { c->err = e; return 1; }
    return 0;
}

}  // extern "C"
