// bam.cu -- BAM input for the native ingest (host code; no device work).
//
// The reference reads BAM through pysam (/root/reference/find_circ.py:461-469).  Here a BAM file (BGZF = concatenated gzip
// members, inflated with zlib) is turned back into SAM text records on the fly -- the eleven mandatory columns plus the
// AS / XS tags, the only optional fields the path looks at (find_circ.py:556-559, 814-817) -- so that BAM input takes the
// same C++ parser, the same fragment logic and the same python fallback as SAM text (csrc/ingest.cu).
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/findcirc_b200.h"

struct fc_bam {
  gzFile gz = nullptr;
  std::vector<std::string> names;
  std::vector<int32_t> lengths;
  std::vector<char> rec;   // current record (binary)
  std::string pending;     // text of a record that did not fit into the caller's buffer
  std::string err;
};

namespace {

bool read_exact(fc_bam* b, void* dst, size_t n) {
  size_t got = 0;
  while (got < n) {
    const int r = gzread(b->gz, (char*)dst + got, (unsigned)((n - got) > (1u << 30) ? (1u << 30) : (n - got)));
    if (r <= 0) return false;
    got += (size_t)r;
  }
  return true;
}

inline int32_t le32(const char* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return (int32_t)v;
}

void append_int(std::string& s, long long v) {
  char buf[24];
  int n = 24;
  unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
  do {
    buf[--n] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  if (v < 0) buf[--n] = '-';
  s.append(buf + n, (size_t)(24 - n));
}

// one binary record -> one SAM text line (appended to `out`); false on a malformed record
bool record_to_text(const fc_bam* b, const char* r, size_t len, std::string& out) {
  if (len < 32) return false;
  const int32_t tid = le32(r), pos = le32(r + 4);
  const uint8_t l_read_name = (uint8_t)r[8], mapq = (uint8_t)r[9];
  uint16_t n_cigar, flag;
  memcpy(&n_cigar, r + 12, 2);
  memcpy(&flag, r + 14, 2);
  const int32_t l_seq = le32(r + 16);
  size_t o = 32;
  if (l_read_name == 0 || l_seq < 0 || o + l_read_name + 4ull * n_cigar + (size_t)(l_seq + 1) / 2 + (size_t)l_seq > len) return false;
  out.append(r + o, (size_t)l_read_name - 1);  // (NUL terminated)
  o += l_read_name;
  out.push_back('\t');
  append_int(out, flag);
  out.push_back('\t');
  if (tid >= 0 && (size_t)tid < b->names.size()) out += b->names[(size_t)tid];
  else out.push_back('*');
  out.push_back('\t');
  append_int(out, (long long)pos + 1);
  out.push_back('\t');
  append_int(out, mapq);
  out.push_back('\t');
  if (n_cigar == 0) {
    out.push_back('*');
  } else {
    for (int k = 0; k < n_cigar; ++k) {
      uint32_t v;
      memcpy(&v, r + o + 4 * k, 4);
      append_int(out, v >> 4);
      const unsigned op = v & 15u;
      if (op > 8) return false;
      out.push_back("MIDNSHP=X"[op]);
    }
  }
  o += 4ull * n_cigar;
  out += "\t*\t0\t0\t";  // mate fields: not looked at on this path
  if (l_seq == 0) {
    out.push_back('*');
  } else {
    static const char code[] = "=ACMGRSVTWYHKDBN";
    const size_t at = out.size();
    out.resize(at + (size_t)l_seq);
    char* d = &out[at];
    const uint8_t* sq = (const uint8_t*)r + o;
    for (int32_t k = 0; k + 1 < l_seq; k += 2) {
      d[k] = code[sq[k >> 1] >> 4];
      d[k + 1] = code[sq[k >> 1] & 15];
    }
    if (l_seq & 1) d[l_seq - 1] = code[sq[l_seq >> 1] >> 4];
  }
  o += (size_t)(l_seq + 1) / 2;
  out.push_back('\t');
  if (l_seq == 0 || (uint8_t)r[o] == 0xFF) {
    out.push_back('*');
  } else {
    const size_t at = out.size();
    out.resize(at + (size_t)l_seq);
    char* d = &out[at];
    for (int32_t k = 0; k < l_seq; ++k) d[k] = (char)((uint8_t)r[o + k] + 33);
  }
  o += (size_t)l_seq;
  // optional fields: AS and XS as integers, everything else is skipped
  while (o + 3 <= len) {
    const char t0 = r[o], t1 = r[o + 1], typ = r[o + 2];
    o += 3;
    long long v = 0;
    bool is_int = true;
    switch (typ) {
      case 'c': if (o + 1 > len) return false; v = (int8_t)r[o]; o += 1; break;
      case 'C': if (o + 1 > len) return false; v = (uint8_t)r[o]; o += 1; break;
      case 's': { if (o + 2 > len) return false; int16_t x; memcpy(&x, r + o, 2); v = x; o += 2; break; }
      case 'S': { if (o + 2 > len) return false; uint16_t x; memcpy(&x, r + o, 2); v = x; o += 2; break; }
      case 'i': { if (o + 4 > len) return false; v = le32(r + o); o += 4; break; }
      case 'I': { if (o + 4 > len) return false; uint32_t x; memcpy(&x, r + o, 4); v = x; o += 4; break; }
      case 'A': is_int = false; o += 1; break;
      case 'f': is_int = false; o += 4; break;
      case 'Z': case 'H': {
        is_int = false;
        const void* e = memchr(r + o, 0, len - o);
        if (!e) return false;
        o = (size_t)((const char*)e - r) + 1;
        break;
      }
      case 'B': {
        is_int = false;
        if (o + 5 > len) return false;
        const char sub = r[o];
        const int32_t cnt = le32(r + o + 1);
        const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
        if (cnt < 0) return false;
        o += 5 + w * (size_t)cnt;
        break;
      }
      default: return false;
    }
    if (o > len) return false;
    if (is_int && ((t0 == 'A' && t1 == 'S') || (t0 == 'X' && t1 == 'S'))) {
      out.push_back('\t');
      out.push_back(t0);
      out.push_back(t1);
      out += ":i:";
      append_int(out, v);
    }
  }
  out.push_back('\n');
  return true;
}

}  // namespace

extern "C" fc_bam* fc_bam_open(const char* path) {
  if (!path) return nullptr;
  fc_bam* b = new fc_bam();
  b->gz = gzopen(path, "rb");
  if (!b->gz) {
    delete b;
    return nullptr;
  }
  gzbuffer(b->gz, 1u << 20);
  char magic[4];
  int32_t l_text = 0, n_ref = 0;
  bool ok = read_exact(b, magic, 4) && memcmp(magic, "BAM\1", 4) == 0 && read_exact(b, &l_text, 4) && l_text >= 0;
  if (ok) {
    std::vector<char> text((size_t)l_text);
    ok = (l_text == 0 || read_exact(b, text.data(), (size_t)l_text)) && read_exact(b, &n_ref, 4) && n_ref >= 0;
  }
  for (int32_t k = 0; ok && k < n_ref; ++k) {
    int32_t l_name = 0, l_ref = 0;
    ok = read_exact(b, &l_name, 4) && l_name > 0;
    if (!ok) break;
    std::string name((size_t)l_name, '\0');
    ok = read_exact(b, &name[0], (size_t)l_name) && read_exact(b, &l_ref, 4);
    name.resize((size_t)l_name - 1);
    b->names.push_back(name);
    b->lengths.push_back(l_ref);
  }
  if (!ok) {
    gzclose(b->gz);
    delete b;
    return nullptr;
  }
  return b;
}

extern "C" void fc_bam_close(fc_bam* b) {
  if (!b) return;
  if (b->gz) gzclose(b->gz);
  delete b;
}

extern "C" int32_t fc_bam_n_ref(const fc_bam* b) { return b ? (int32_t)b->names.size() : -1; }
extern "C" const char* fc_bam_ref_name(const fc_bam* b, int32_t i) {
  return (b && i >= 0 && (size_t)i < b->names.size()) ? b->names[(size_t)i].c_str() : nullptr;
}
extern "C" int64_t fc_bam_ref_length(const fc_bam* b, int32_t i) {
  return (b && i >= 0 && (size_t)i < b->lengths.size()) ? b->lengths[(size_t)i] : -1;
}

// Fills `out` with whole SAM text lines of the next records (at most `cap` bytes, cap >= 64 KiB).  Returns the bytes
// written, 0 at the end of the file, FC_E_IO for a truncated or malformed file.
extern "C" int64_t fc_bam_read_text(fc_bam* b, char* out, int64_t cap) {
  if (!b || !out || cap < (1 << 16)) return FC_E_ARG;
  int64_t w = 0;
  std::string line;
  for (;;) {
    if (!b->pending.empty()) {
      if ((int64_t)b->pending.size() > cap - w) {
        if (w == 0) return FC_E_ARG;  // one record larger than the whole buffer
        return w;
      }
      memcpy(out + w, b->pending.data(), b->pending.size());
      w += (int64_t)b->pending.size();
      b->pending.clear();
    }
    int32_t block = 0;
    const int r = gzread(b->gz, &block, 4);
    if (r == 0) return w;  // clean end of file
    if (r != 4 || block < 32) return FC_E_IO;
    b->rec.resize((size_t)block);
    if (!read_exact(b, b->rec.data(), (size_t)block)) return FC_E_IO;
    line.clear();
    if (!record_to_text(b, b->rec.data(), (size_t)block, line)) return FC_E_IO;
    b->pending.swap(line);
    if (cap - w < (1 << 12) && (int64_t)b->pending.size() > cap - w) return w;
  }
}
