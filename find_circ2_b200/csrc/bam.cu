// bam.cu -- BAM input for the native ingest (host code; no device work).
//
// The reference reads BAM through pysam (/root/reference/find_circ.py:461-469).  Here a BAM file (BGZF = concatenated gzip
// members, inflated with zlib) is turned back into SAM text records on the fly -- the eleven mandatory columns plus the
// AS / XS tags, the only optional fields the path looks at (find_circ.py:556-559, 814-817) -- so that BAM input takes the
// same C++ parser, the same fragment logic and the same python fallback as SAM text (csrc/ingest.cu).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/findcirc_b200.h"

struct fc_bam {
  gzFile gz = nullptr;     // files that are gzip but not BGZF (no block sizes in the headers): one inflate stream
  FILE* fp = nullptr;      // BGZF: blocks are read raw and inflated on several threads
  std::vector<std::string> names;
  std::vector<int32_t> lengths;
  std::vector<char> rec;   // current record (binary)
  std::string pending;     // text of a record that did not fit into the caller's buffer
  std::string err;
  // BGZF state: inflated bytes not yet turned into text, and text not yet handed out
  std::vector<char> ubuf;
  size_t upos = 0;
  std::vector<std::string> ready;
  size_t ready_idx = 0, ready_off = 0;
  bool at_eof = false, broken = false;
  int threads = 1;
  std::vector<unsigned char> comp;  // the compressed members of the current batch (kept: no fresh pages per batch)
};

namespace {

bool bgzf_read(fc_bam* b, void* dst, size_t n);
bool read_exact(fc_bam* b, void* dst, size_t n) {
  if (b->fp) return bgzf_read(b, dst, n);
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

// ---- BGZF on several threads --------------------------------------------------------------------------------------
// A BGZF file is a series of gzip members of at most 64 KiB whose headers carry the member's size ("BC" extra field), so
// the members can be found without inflating and inflated independently.  bgzf_refill reads a batch of members, inflates
// them on b->threads threads into one buffer (behind what the last batch left unconsumed: a record may straddle
// batches) -- the reference does all of this inside pysam on one thread (find_circ.py:461-469).
constexpr size_t BGZF_BATCH = 512;  // members per batch (<= 32 MiB inflated)

struct Member {
  size_t coff, csize;  // in the batch's compressed buffer: the whole member
  size_t hdr;          // length of its header
  uint32_t isize, crc;
  size_t uoff;         // where its bytes go in ubuf
};

// true: *size = total size of the gzip member that starts with the 12 + xlen header bytes at p (n bytes available)
bool bgzf_member_size(const unsigned char* p, size_t n, size_t* size, size_t* hdr) {
  if (n < 12 || p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return false;
  const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
  if (n < 12 + xlen) return false;
  for (size_t o = 12; o + 4 <= 12 + xlen;) {
    const size_t slen = (size_t)p[o + 2] | ((size_t)p[o + 3] << 8);
    if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) {
      *size = ((size_t)p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
      *hdr = 12 + xlen;
      return *size >= *hdr + 8;
    }
    o += 4 + slen;
  }
  return false;
}

bool inflate_member(const unsigned char* src, const Member& m, char* dst) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<unsigned char*>(src + m.coff + m.hdr);
  zs.avail_in = (uInt)(m.csize - m.hdr - 8);
  zs.next_out = reinterpret_cast<unsigned char*>(dst);
  zs.avail_out = m.isize;
  const int rc = m.isize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
  const bool ok = (rc == Z_STREAM_END || (m.isize == 0 && rc == Z_OK)) && zs.avail_out == 0;
  inflateEnd(&zs);
  return ok && (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char*>(dst), m.isize) == m.crc;
}

// appends the next batch of members to ubuf (after dropping what has been consumed); false: error (b->broken) or nothing
// left (b->at_eof)
bool bgzf_refill(fc_bam* b) {
  if (b->at_eof || b->broken) return false;
  if (b->upos) {
    b->ubuf.erase(b->ubuf.begin(), b->ubuf.begin() + (ptrdiff_t)b->upos);
    b->upos = 0;
  }
  std::vector<unsigned char>& comp = b->comp;
  comp.clear();
  std::vector<Member> ms;
  size_t total = b->ubuf.size();
  while (ms.size() < BGZF_BATCH) {
    unsigned char head[12];
    const size_t got = fread(head, 1, 12, b->fp);
    if (got == 0) {
      b->at_eof = true;
      break;
    }
    if (got != 12) { b->broken = true; return false; }
    const size_t xlen = (size_t)head[10] | ((size_t)head[11] << 8);
    const size_t at = comp.size();
    comp.resize(at + 12 + xlen);
    memcpy(comp.data() + at, head, 12);
    if (fread(comp.data() + at + 12, 1, xlen, b->fp) != xlen) { b->broken = true; return false; }
    size_t size = 0, hdr = 0;
    if (!bgzf_member_size(comp.data() + at, 12 + xlen, &size, &hdr)) { b->broken = true; return false; }
    comp.resize(at + size);
    if (fread(comp.data() + at + hdr, 1, size - hdr, b->fp) != size - hdr) { b->broken = true; return false; }
    Member m;
    m.coff = at;
    m.csize = size;
    m.hdr = hdr;
    memcpy(&m.crc, comp.data() + at + size - 8, 4);
    memcpy(&m.isize, comp.data() + at + size - 4, 4);
    if (m.isize > (1u << 16)) { b->broken = true; return false; }
    m.uoff = total;
    total += m.isize;
    ms.push_back(m);
  }
  if (ms.empty()) return false;
  b->ubuf.resize(total);
  const int T = (int)(ms.size() < (size_t)b->threads ? ms.size() : (size_t)b->threads);
  std::vector<char> ok((size_t)T, 1);
  auto work = [&](int t) {
    for (size_t k = (size_t)t; k < ms.size(); k += (size_t)T)
      if (!inflate_member(comp.data(), ms[k], b->ubuf.data() + ms[k].uoff)) ok[(size_t)t] = 0;
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  for (char o : ok)
    if (!o) { b->broken = true; return false; }
  return true;
}

// at least n inflated bytes from upos on? (refills as needed)
bool bgzf_need(fc_bam* b, size_t n) {
  while (b->ubuf.size() - b->upos < n)
    if (!bgzf_refill(b)) return false;
  return true;
}
bool bgzf_read(fc_bam* b, void* dst, size_t n) {
  if (!bgzf_need(b, n)) return false;
  memcpy(dst, b->ubuf.data() + b->upos, n);
  b->upos += n;
  return true;
}

// the whole records in ubuf as text, formatted on b->threads threads into b->ready (one piece per thread, in order);
// false at the end of the file (or on an error: b->broken)
bool bgzf_format_batch(fc_bam* b) {
  for (;;) {
    std::vector<size_t> offs;
    size_t p = b->upos;
    while (b->ubuf.size() - p >= 4) {
      const int32_t block = le32(b->ubuf.data() + p);
      if (block < 32) { b->broken = true; return false; }
      if (b->ubuf.size() - p - 4 < (size_t)block) break;
      offs.push_back(p);
      p += 4 + (size_t)block;
    }
    if (offs.empty()) {
      if (bgzf_refill(b)) continue;
      if (!b->broken && b->ubuf.size() != b->upos) b->broken = true;  // the file ends inside a record
      return false;
    }
    const int T = (int)(offs.size() < (size_t)b->threads * 64 ? 1 : b->threads);
    b->ready.resize((size_t)T);  // (the strings keep their memory from batch to batch)
    for (std::string& piece : b->ready) piece.clear();
    b->ready_idx = b->ready_off = 0;
    std::vector<char> ok((size_t)T, 1);
    auto work = [&](int t) {
      const size_t k0 = offs.size() * (size_t)t / (size_t)T, k1 = offs.size() * (size_t)(t + 1) / (size_t)T;
      std::string& out = b->ready[(size_t)t];
      for (size_t k = k0; k < k1; ++k) {
        const char* r = b->ubuf.data() + offs[k];
        if (!record_to_text(b, r + 4, (size_t)le32(r), out)) { ok[(size_t)t] = 0; return; }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    b->upos = p;
    for (char o : ok)
      if (!o) { b->broken = true; return false; }
    return true;
  }
}

}  // namespace

extern "C" fc_bam* fc_bam_open(const char* path) {
  if (!path) return nullptr;
  fc_bam* b = new fc_bam();
  {
    // BGZF (what every BAM writer produces)?  Then the members are inflated on several threads; any other gzip file goes
    // through one zlib stream
    FILE* fp = fopen(path, "rb");
    if (!fp) {
      delete b;
      return nullptr;
    }
    unsigned char head[64];
    const size_t got = fread(head, 1, sizeof(head), fp);
    size_t size = 0, hdr = 0;
    if (bgzf_member_size(head, got, &size, &hdr)) {
      fseek(fp, 0, SEEK_SET);
      b->fp = fp;
      unsigned hc = std::thread::hardware_concurrency();
      b->threads = hc == 0 ? 1 : (hc > 16 ? 16 : (int)hc);
      if (const char* e = getenv("FC_BAM_THREADS")) b->threads = atoi(e) > 0 ? atoi(e) : b->threads;
    } else {
      fclose(fp);
    }
  }
  if (!b->fp) {
    b->gz = gzopen(path, "rb");
    if (!b->gz) {
      delete b;
      return nullptr;
    }
    gzbuffer(b->gz, 1u << 20);
  }
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
    if (b->gz) gzclose(b->gz);
    if (b->fp) fclose(b->fp);
    delete b;
    return nullptr;
  }
  return b;
}

extern "C" void fc_bam_close(fc_bam* b) {
  if (!b) return;
  if (b->gz) gzclose(b->gz);
  if (b->fp) fclose(b->fp);
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
  if (b->fp) {
    // pieces of text formatted batch-wise on several threads; handed out in whole lines
    for (;;) {
      while (b->ready_idx < b->ready.size()) {
        const std::string& piece = b->ready[b->ready_idx];
        size_t left = piece.size() - b->ready_off;
        if (left == 0) {
          b->ready_idx++;
          b->ready_off = 0;
          continue;
        }
        size_t take = left;
        if ((int64_t)take > cap - w) {
          take = (size_t)(cap - w);
          while (take > 0 && piece[b->ready_off + take - 1] != '\n') --take;  // (whole lines only)
          if (take == 0) {
            if (w == 0) return FC_E_ARG;  // one record larger than the whole buffer
            return w;
          }
        }
        memcpy(out + w, piece.data() + b->ready_off, take);
        w += (int64_t)take;
        b->ready_off += take;
        if (take < left) return w;
      }
      if (cap - w < (1 << 12)) return w;
      if (!bgzf_format_batch(b)) return b->broken ? (int64_t)FC_E_IO : w;
    }
  }
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
