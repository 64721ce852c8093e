// ingest.cu -- native SAM-text ingest for the common fragment shapes (host code; no device work).
//
// The reference decodes every record through pysam and builds Python objects per read
// (/root/reference/find_circ.py:461-469, 1450-1486, 976-1140, 821-852); a python host that does the same is three orders
// of magnitude slower than the scan kernel.  This file parses SAM text in C++, groups the records into mates and
// fragments, forms the anchor pairs ("spans") of every mate and emits, for every fragment with at most two spans in
// total (single-end two-segment reads, mate pairs with one or both mates spliced, three-segment mates: the bulk of real
// input), the struct-of-arrays rows the GPU needs -- window coordinates, flags, the internal read part already as bit
// planes, anchor qualities and the read / name hashes, ready to be copied to the device as they are -- plus one record
// per fragment with what the evidence rules of record_hits (find_circ.py:1276-1439) need besides the scan's answers
// (pipeline.py applies them to whole batches at once).
// Everything else (three or more spans, a third mate, records it cannot interpret) is handed back as a byte range and
// goes through the python implementation of the same logic (find_circ2_b200/pipeline.py), so the two paths together cover
// exactly what the reference covers.  Counters are accumulated here only for the fragments handled here.
#include <emmintrin.h>
#include <string.h>

#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "fc_internal.cuh"

extern "C" uint64_t fc_hash_bytes(const uint8_t* p, int64_t n);
extern "C" uint64_t fc_hash_read(const uint8_t* seq, int64_t n, int32_t* is_palindrome);

namespace {

using sv = std::string_view;

struct Rec {
  int64_t off = 0, len = 0;  // line in the chunk
  sv qname, seq, qual;
  int flag = 0, tid = -1, pos = 0, aend = 0, clip_start = 0, qlen = 0, AS = 0, XS = 0;
  bool has_cigar = false, has_AS = false, has_XS = false, has_seq = false, has_qual = false, ok = true;
};

inline bool parse_int(sv s, int& out) {
  if (s.empty()) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '-') { neg = true; i = 1; }
  else if (s[0] == '+') i = 1;
  if (i >= s.size()) return false;
  long v = 0;
  for (; i < s.size(); ++i) {
    char c = s[i];
    if (c < '0' || c > '9') return false;
    v = v * 10 + (c - '0');
    if (v > 2147483647L) return false;
  }
  out = (int)(neg ? -v : v);
  return true;
}

// positions of the first `cap` tab characters of a line
inline int find_tabs(const char* p, size_t n, uint32_t* out, int cap) {
  int k = 0;
  size_t i = 0;
  const __m128i tabs = _mm_set1_epi8('\t');
  for (; i + 16 <= n && k < cap; i += 16) {
    unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(p + i)), tabs));
    while (m && k < cap) {
      out[k++] = (uint32_t)(i + (unsigned)__builtin_ctz(m));
      m &= m - 1u;
    }
  }
  for (; i < n && k < cap; ++i)
    if (p[i] == '\t') out[k++] = (uint32_t)i;
  return k;
}

struct Ingest {
  fc_ingest_params p;
  std::unordered_map<std::string, int> name2tid;
  std::vector<int32_t> tid2gid;
  std::string last_name;
  int last_tid = -1;
  int64_t frag_seq = 0;
  bool first_record = true;
  int64_t planes_rows = -1;  // rows / words of the plane arrays that the last call has written (-1: unknown, clear everything)
  int planes_words = 0;
  const void* planes_of = nullptr;

  int lookup(sv name) {
    if (name.size() == last_name.size() && memcmp(name.data(), last_name.data(), name.size()) == 0) return last_tid;
    auto it = name2tid.find(std::string(name));
    last_name.assign(name.data(), name.size());
    last_tid = it == name2tid.end() ? -1 : it->second;
    return last_tid;
  }

  // pysam field semantics (find_circ2_b200/samio.py is the python twin of this function)
  void parse_line(const char* base, int64_t off, int64_t len, Rec& r) {
    r = Rec();
    r.off = off;
    r.len = len;
    sv line(base + off, (size_t)len);
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.remove_suffix(1);
    // all tabs of the line in one sweep (16 bytes per step): a record has 11 mandatory columns and a few tags
    uint32_t tab[48];
    const int n_tabs = find_tabs(line.data(), line.size(), tab, 48);
    if (n_tabs < 10) { r.ok = false; return; }
    sv f[11];
    size_t start = 0;
    for (int k = 0; k < 11; ++k) {
      const size_t t = k < n_tabs ? tab[k] : line.size();
      f[k] = line.substr(start, t - start);
      start = t < line.size() ? t + 1 : line.size();
    }
    r.qname = f[0];
    if (!parse_int(f[1], r.flag)) { r.ok = false; return; }
    r.tid = f[2] == "*" ? -1 : lookup(f[2]);
    int pos1;
    if (!parse_int(f[3], pos1)) { r.ok = false; return; }
    r.pos = pos1 - 1;
    // CIGAR
    if (f[5] != "*") {
      r.has_cigar = true;
      int aend = r.pos, clip = 0, lead_soft = 0, trail_soft = 0;
      bool in_lead = true, clip_done = false, any = false;
      long num = 0;
      bool have_num = false;
      int last_soft = 0;  // trailing soft clip candidate
      for (char c : f[5]) {
        if (c >= '0' && c <= '9') { num = num * 10 + (c - '0'); have_num = true; continue; }
        if (!have_num) { r.ok = false; return; }
        int n = (int)num;
        num = 0;
        have_num = false;
        any = true;
        switch (c) {
          case 'M': case '=': case 'X': aend += n; break;
          case 'D': case 'N': aend += n; break;
          case 'I': case 'P': break;
          case 'S': case 'H': break;
          default: r.ok = false; return;
        }
        // offset of the aligned part: S/H lengths are summed until the first M; other operations do not stop the walk
        // (find_circ.py:1086-1097)
        if (!clip_done) {
          if (c == 'S' || c == 'H') clip += n;
          else if (c == 'M') clip_done = true;
        }
        // soft clips at the two ends of the stored sequence (hard clips may sit outside them)
        if (in_lead) {
          if (c == 'H') { /* skip */ }
          else if (c == 'S') lead_soft += n;
          else in_lead = false;
        }
        if (c == 'S') last_soft += n;
        else if (c != 'H') last_soft = 0;
      }
      if (have_num || !any) { r.ok = false; return; }
      trail_soft = in_lead ? 0 : last_soft;  // an all-clip CIGAR has no trailing part of its own
      r.aend = aend;
      r.clip_start = clip;
      r.qlen = -(lead_soft + trail_soft);  // completed below with the sequence length
    }
    if (f[9] != "*") { r.has_seq = true; r.seq = f[9]; }
    if (f[10] != "*") { r.has_qual = true; r.qual = f[10]; }
    r.qlen += (int)r.seq.size();
    // tags
    int next_tab = 11;
    while (start < line.size()) {
      size_t t;
      if (next_tab < n_tabs) t = tab[next_tab++];
      else if (n_tabs < 48) t = sv::npos;
      else t = line.find('\t', start);  // (more tabs than the sweep recorded)
      sv tag = line.substr(start, t == sv::npos ? sv::npos : t - start);
      if (tag.size() > 5 && tag[2] == ':' && tag[3] == 'i' && tag[4] == ':') {
        if (tag[0] == 'A' && tag[1] == 'S') r.has_AS = parse_int(tag.substr(5), r.AS);
        else if (tag[0] == 'X' && tag[1] == 'S') r.has_XS = parse_int(tag.substr(5), r.XS);
      }
      if (t == sv::npos) break;
      start = t + 1;
    }
  }
};

// 8 bases -> 8 bits of the low plane, the high plane and the N plane (bit j = base j).  Upper-cased ASCII: A 0x41, C 0x43,
// G 0x47, T 0x54, so code bit 0 = bit 1 ^ bit 2 and code bit 1 = bit 2 of the letter; anything else is N (code 0), like
// fc::pack32 (scan_core.cuh), which the device uses and tests/ compare with.
inline uint64_t zero_bytes(uint64_t v) {  // 0x80 in every byte of v that is zero (exact)
  const uint64_t k = 0x7F7F7F7F7F7F7F7FULL;
  return ~(((v & k) + k) | v | k);
}
inline uint32_t gather_bit0(uint64_t v) {  // bit 0 of every byte -> 8 bits
  return (uint32_t)(((v & 0x0101010101010101ULL) * 0x0102040810204080ULL) >> 56);
}
inline void pack8(uint64_t x, uint32_t& lo, uint32_t& hi, uint32_t& nn) {
  x &= 0xDFDFDFDFDFDFDFDFULL;
  const uint64_t acgt = zero_bytes(x ^ 0x4141414141414141ULL) | zero_bytes(x ^ 0x4343434343434343ULL) |
                        zero_bytes(x ^ 0x4747474747474747ULL) | zero_bytes(x ^ 0x5454545454545454ULL);
  uint64_t l = (x >> 1) ^ (x >> 2), h = x >> 2;
  if (acgt != 0x8080808080808080ULL) {
    const uint64_t keep = (acgt >> 7) * 0xFFULL;  // 0xFF in the bytes that hold A, C, G or T
    l &= keep;
    h &= keep;
    nn = gather_bit0(~acgt >> 7);
  } else {
    nn = 0;
  }
  lo = gather_bit0(l);
  hi = gather_bit0(h);
}
inline void pack32_host(const uint8_t* src, int count, uint32_t& lo, uint32_t& hi, uint32_t& nn) {
  lo = hi = nn = 0;
  for (int j = 0; j < count; j += 8) {
    uint64_t x = 0x4141414141414141ULL;  // (bases beyond the end read as A = all planes 0)
    const int take = count - j < 8 ? count - j : 8;
    memcpy(&x, src + j, (size_t)take);
    uint32_t a, b, c;
    pack8(x, a, b, c);
    lo |= a << j;
    hi |= b << j;
    nn |= c << j;
  }
}

// the internal read part of row `row` as bit planes (column-major: word w of every row is contiguous).  Only the words the
// part reaches are written: fc_ingest_parse clears what an earlier call left in the arrays, in bulk.
inline int pack_planes(sv s, int n_words, int64_t stride, int64_t row, uint32_t* rlo, uint32_t* rhi, uint32_t* rn, bool& any_n) {
  const int used = ((int)s.size() + 31) / 32 < n_words ? ((int)s.size() + 31) / 32 : n_words;
  for (int w = 0; w < used; ++w) {
    uint32_t lo = 0, hi = 0, nn = 0;
    int count = (int)s.size() - 32 * w;
    if (count > 32) count = 32;
    pack32_host(reinterpret_cast<const uint8_t*>(s.data()) + 32 * w, count, lo, hi, nn);
    rlo[(int64_t)w * stride + row] = lo;
    rhi[(int64_t)w * stride + row] = hi;
    rn[(int64_t)w * stride + row] = nn;
    if (nn) any_n = true;
  }
  return used;
}

}  // namespace

struct fc_ingest {
  Ingest g;
};

extern "C" fc_ingest* fc_ingest_create(const fc_ingest_params* p, int32_t n_names, const char* const* names,
                                       const int32_t* tid2gid) {
  if (!p || n_names < 0) return nullptr;
  fc_ingest* h = new fc_ingest();
  h->g.p = *p;
  for (int32_t i = 0; i < n_names; ++i) {
    h->g.name2tid.emplace(names[i], i);
    h->g.tid2gid.push_back(tid2gid ? tid2gid[i] : i);
  }
  return h;
}

extern "C" void fc_ingest_destroy(fc_ingest* h) { delete h; }

// a parser that starts in the middle of the input stream (one rank of a multi-GPU run): ordinal of its first fragment,
// and whether its first record is the first record of the whole stream (which the reference never checks for the
// "unmapped" flag, find_circ.py:1462-1463)
extern "C" int64_t fc_ingest_position(fc_ingest* h) { return h ? h->g.frag_seq : -1; }  // ordinal the next fragment will get

extern "C" int fc_ingest_set_position(fc_ingest* h, int64_t first_fragment, int32_t at_stream_start) {
  if (!h || first_fragment < 0) return FC_E_ARG;
  h->g.frag_seq = first_fragment;
  h->g.first_record = at_stream_start != 0;
  return FC_OK;
}

extern "C" int64_t fc_ingest_parse(fc_ingest* h, const char* text, int64_t nbytes, int32_t final, fc_ingest_out* o) {
  if (!h || !text || !o || nbytes < 0) return FC_E_ARG;
  Ingest& g = h->g;
  const fc_ingest_params& P = g.p;
  const int eff = P.asize - P.margin;
  o->n_rows = 0;
  o->n_frag_records = 0;
  o->n_complex = 0;
  o->n_fragments = 0;
  o->max_l = 0;
  for (int k = 0; k < 8; ++k) o->counters[k] = 0;
  {
    // rows write only the plane words their read part reaches: whatever an earlier call left behind goes now, in bulk
    // (the first call, or a call with other arrays, clears them whole)
    const bool same = g.planes_of == (const void*)o->rlo && g.planes_rows >= 0;
    const int words = same ? g.planes_words : o->n_words;
    const int64_t rows = same ? g.planes_rows : o->cap;
    for (int w = 0; w < words; ++w) {
      memset(o->rlo + (int64_t)w * o->cap, 0, (size_t)rows * 4);
      memset(o->rhi + (int64_t)w * o->cap, 0, (size_t)rows * 4);
      memset(o->rn + (int64_t)w * o->cap, 0, (size_t)rows * 4);
    }
    g.planes_of = o->rlo;
    g.planes_rows = 0;
    g.planes_words = 0;
  }
  enum { C_TOTAL_MATES, C_UNMAPPED, C_UNSPLICED, C_TOO_SHORT, C_CIRC_NOT_UNIQ, C_LIN_NOT_UNIQ };

  std::vector<Rec> frag;       // mapped records of the current fragment (first record regardless of its flag)
  frag.reserve(8);
  int64_t frag_start = 0;      // byte offset of the first line of the current fragment
  int64_t frag_unmapped = 0;   // unmapped records skipped inside the current fragment's byte range
  bool force_python = false;   // a line the parser could not interpret lies in or next to the fragment
  bool have_frag = false;
  int64_t consumed = 0;

  auto leave_to_python = [&](int64_t frag_end) -> bool {
    if (o->n_complex >= o->cap_complex) return false;
    o->cx_start[o->n_complex] = frag_start;
    o->cx_end[o->n_complex] = frag_end;
    o->cx_seq[o->n_complex] = g.frag_seq;
    o->n_complex++;
    o->n_fragments++;
    g.frag_seq++;
    return true;
  };

  // one anchor pair of a mate (JunctionSpan, find_circ.py:821-852)
  struct SpanC {
    const Rec *A, *B, *primary;
    int q_start, q_end, den;
    bool backsplice, queued;
  };
  constexpr int MAX_REC = 16;

  auto finish = [&](int64_t frag_end) -> bool {
    // returns false when the output arrays are full (the fragment is then NOT consumed)
    const int nrec = (int)frag.size();
    bool complex = force_python || nrec > MAX_REC || (frag[0].flag & 0x4) != 0;
    for (const Rec& r : frag) complex = complex || !r.ok;
    if (complex) return leave_to_python(frag_end);
    double add[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // ---- mates (find_circ.py:1450-1486): a record whose read1 flag differs from the current mate's primary starts the
    // other mate; a third mate would push the first one out -- left to python
    int mate_of[MAX_REC], prim[2] = {0, 0}, n_mates = 1;
    mate_of[0] = 0;
    for (int k = 1; k < nrec; ++k) {
      if (((frag[k].flag ^ frag[prim[n_mates - 1]].flag) & 0x40) != 0) {
        if (n_mates == 2) return leave_to_python(frag_end);
        prim[n_mates++] = k;
      }
      mate_of[k] = n_mates - 1;
    }
    add[C_TOTAL_MATES] += n_mates;
    add[C_UNMAPPED] += (double)frag_unmapped;
    // ---- segments and spans of every mate (MateSegments / adjacent_segment_pairs, find_circ.py:976-1140)
    SpanC circ[3], lin[3];
    int n_circ = 0, n_lin = 0;
    const Rec* unspliced = nullptr;
    bool broken = false;
    for (int m = 0; m < n_mates; ++m) {
      const Rec& p = frag[prim[m]];
      int proper[MAX_REC], n_proper = 0, n_other = 0;
      for (int k = 0; k < nrec; ++k) {
        if (mate_of[k] != m) continue;
        const Rec& r = frag[k];
        if (r.tid == p.tid && ((r.flag ^ p.flag) & 0x10) == 0) proper[n_proper++] = k;
        else ++n_other;
      }
      if (n_proper < 2) {
        add[C_UNSPLICED] += 1;
        if (!unspliced) unspliced = &p;
        continue;
      }
      if (n_proper - 1 > 255) return leave_to_python(frag_end);  // (python raises on these)
      for (int k = 0; k < n_proper; ++k) {
        const Rec& r = frag[proper[k]];
        if (!r.has_seq || !r.has_cigar) return leave_to_python(frag_end);  // python raises on these, keep its behaviour
      }
      // stable order by the start of the aligned part in the read
      for (int i = 1; i < n_proper; ++i) {
        const int v = proper[i];
        int j = i - 1;
        while (j >= 0 && frag[proper[j]].clip_start > frag[v].clip_start) { proper[j + 1] = proper[j]; --j; }
        proper[j + 1] = v;
      }
      const int L = (int)p.seq.size();
      int lo = L, hi = 0;
      for (int k = 0; k + 1 < n_proper; ++k) {
        const Rec *a = &frag[proper[k]], *b = &frag[proper[k + 1]];
        if (a->qlen < P.asize || b->qlen < P.asize) {
          add[C_TOO_SHORT] += 1;
          continue;
        }
        SpanC sp;
        sp.A = a;
        sp.B = b;
        sp.primary = &p;
        sp.q_start = a->clip_start < b->clip_start ? a->clip_start : b->clip_start;
        const int ea = a->clip_start + a->qlen, eb = b->clip_start + b->qlen;
        sp.q_end = ea > eb ? ea : eb;
        sp.den = n_proper - 1;
        sp.backsplice = (b->pos - a->aend) < 0;
        sp.queued = false;
        if (!a->has_AS || !b->has_AS) return leave_to_python(frag_end);
        if (sp.q_start < lo) lo = sp.q_start;
        if (sp.q_end > hi) hi = sp.q_end;
        if (sp.backsplice) {
          if (n_circ == 3) return leave_to_python(frag_end);
          circ[n_circ++] = sp;
        } else {
          if (n_lin == 3) return leave_to_python(frag_end);
          lin[n_lin++] = sp;
        }
      }
      if ((hi < L - P.asize || lo > P.asize) && n_other > 0) broken = true;
    }
    // ---- the head of record_hits (find_circ.py:1560-1574): which spans go to the scan
    const bool dropped = (n_circ == 0 && P.nolinear) || (n_circ + n_lin == 0);
    int n_rows_new = 0;
    SpanC* spans[2] = {nullptr, nullptr};
    int n_sp = 0;
    if (!dropped) {
      if (n_circ + n_lin > 2 || (P.nolinear && n_circ > 0 && n_lin > 0)) return leave_to_python(frag_end);
      for (int k = 0; k < n_circ; ++k) spans[n_sp++] = &circ[k];
      for (int k = 0; k < n_lin; ++k) spans[n_sp++] = &lin[k];
      for (int k = 0; k < n_sp; ++k) {
        SpanC& sp = *spans[k];
        const int ua = sp.A->has_XS ? sp.A->AS - sp.A->XS : sp.A->AS, ub = sp.B->has_XS ? sp.B->AS - sp.B->XS : sp.B->AS;
        if ((ua < ub ? ua : ub) < P.min_uniq_qual) {
          add[sp.backsplice ? C_CIRC_NOT_UNIQ : C_LIN_NOT_UNIQ] += 1;
          continue;
        }
        const int tid = sp.A->tid;
        if (tid < 0 || tid >= (int)g.tid2gid.size() || g.tid2gid[tid] < 0) return leave_to_python(frag_end);  // (python raises)
        const int plen = (int)sp.primary->seq.size();
        int qs = sp.q_start < plen ? sp.q_start : plen, qe = sp.q_end < plen ? sp.q_end : plen;
        if (qe < qs) qe = qs;
        const int l = qe - qs - 2 * eff;
        // read longer than the caller's plane buffer: let the python path deal with it
        if (l > 0 && (l + 31) / 32 > o->n_words) return leave_to_python(frag_end);
        sp.queued = true;
        ++n_rows_new;
      }
    }
    if (n_rows_new) {
      for (int m = 0; m < n_mates; ++m)
        if (!frag[prim[m]].has_seq) return leave_to_python(frag_end);  // (python fails when it prints such a read)
      if (o->n_rows + n_rows_new > o->cap) return false;
      const int64_t fi = o->n_frag_records;
      o->f_seq[fi] = g.frag_seq;
      o->f_row0[fi] = (int32_t)o->n_rows;
      o->f_nsp[fi] = (uint8_t)n_sp;
      uint8_t kind = 0, state = 0;
      int kq = 0;
      for (int k = 0; k < n_sp; ++k) {
        const SpanC& sp = *spans[k];
        if (sp.backsplice) kind |= (uint8_t)(1u << k);
        if (!sp.queued) continue;
        state |= (uint8_t)(1u << k);
        const int64_t i = o->n_rows;
        const Rec& p = *sp.primary;
        const int plen = (int)p.seq.size();
        int qs = sp.q_start < plen ? sp.q_start : plen, qe = sp.q_end < plen ? sp.q_end : plen;
        if (qe < qs) qe = qs;
        const int L = qe - qs;
        const int l = L - 2 * eff;
        sv internal;
        if (L - eff > eff) internal = p.seq.substr((size_t)(qs + eff), (size_t)(L - 2 * eff));
        o->chrom[i] = g.tid2gid[sp.A->tid];
        o->a_start[i] = sp.A->pos + eff;
        o->b_end[i] = sp.B->aend - eff;
        o->l[i] = l;
        bool any_n = false;
        const int used = pack_planes(internal, o->n_words, o->cap, i, o->rlo, o->rhi, o->rn, any_n);
        if (used > g.planes_words) g.planes_words = used;
        o->flags[i] = (uint8_t)((sp.backsplice ? 1 : 0) | ((p.flag & 0x10) ? 2 : 0) | (any_n ? 4 : 0));
        o->wden[i] = (uint8_t)sp.den;
        const int qa = sp.A->AS - (sp.A->has_XS ? sp.A->XS : 0), qb = sp.B->AS - (sp.B->has_XS ? sp.B->XS : 0);
        o->q_a[i] = (int16_t)(qa < -32768 ? -32768 : (qa > 32767 ? 32767 : qa));
        o->q_b[i] = (int16_t)(qb < -32768 ? -32768 : (qb > 32767 ? 32767 : qb));
        o->read_hash[i] = fc_hash_read(reinterpret_cast<const uint8_t*>(p.seq.data()), (int64_t)p.seq.size(), nullptr);
        o->qname_hash[i] = fc_hash_bytes(reinterpret_cast<const uint8_t*>(p.qname.data()), (int64_t)p.qname.size());
        o->frag_seq[i] = g.frag_seq;
        o->idx_k[i] = (uint8_t)kq++;
        if (l > o->max_l) o->max_l = l;
        o->n_rows++;
        g.planes_rows = o->n_rows;
      }
      o->f_kind[fi] = kind;
      o->f_state[fi] = state;
      uint8_t ff = (uint8_t)((n_mates == 2 ? FC_FR_TWO_MATES : 0u) | (broken ? FC_FR_BROKEN : 0u));
      o->f_un_tid[fi] = o->f_un_pos[fi] = o->f_un_aend[fi] = 0;
      if (unspliced) {
        ff |= FC_FR_UNSPLICED;
        if (n_circ > 0 && circ[0].primary->tid != unspliced->tid) ff |= FC_FR_OTHER_CHROM;
        o->f_un_tid[fi] = unspliced->tid;
        o->f_un_pos[fi] = unspliced->pos;
        o->f_un_aend[fi] = unspliced->aend;
      }
      o->f_flags[fi] = ff;
      for (int m = 0; m < 2; ++m) {
        int64_t* off = o->f_txt_off + (fi * 2 + m) * 3;
        int32_t* len = o->f_txt_len + (fi * 2 + m) * 3;
        if (m >= n_mates) {
          off[0] = off[1] = off[2] = 0;
          len[0] = len[1] = len[2] = -2;
          continue;
        }
        const Rec& p = frag[prim[m]];
        off[0] = p.qname.data() - text;
        len[0] = (int32_t)p.qname.size();
        off[1] = p.has_seq ? p.seq.data() - text : 0;
        len[1] = p.has_seq ? (int32_t)p.seq.size() : -1;
        off[2] = p.has_qual ? p.qual.data() - text : 0;
        len[2] = p.has_qual ? (int32_t)p.qual.size() : -1;
      }
      o->n_frag_records++;
    }
    for (int k = 0; k < 8; ++k) o->counters[k] += add[k];
    o->n_fragments++;
    g.frag_seq++;
    return true;
  };

  int64_t pos = 0;
  Rec r;
  while (pos < nbytes) {
    const char* nl = (const char*)memchr(text + pos, '\n', (size_t)(nbytes - pos));
    int64_t line_end;
    if (!nl) {
      if (!final) break;  // incomplete last line: wait for more text
      line_end = nbytes;
    } else {
      line_end = (nl - text) + 1;
    }
    const int64_t line_off = pos;
    const int64_t line_len = line_end - pos;
    pos = line_end;
    // skip header and blank lines
    if (text[line_off] == '@' || line_len <= 1 || (line_len == 2 && text[line_off] == '\r')) {
      if (!have_frag) consumed = pos;
      continue;
    }
    g.parse_line(text, line_off, line_len, r);
    if (!have_frag) {
      if (!g.first_record && r.ok && (r.flag & 0x4)) {
        // an unmapped record where a chunk (or a rank's part of the stream) begins: counted and skipped like any other;
        // only the very first record of the stream escapes that check (find_circ.py:1462-1467)
        o->counters[C_UNMAPPED] += 1;
        consumed = pos;
        continue;
      }
      frag.clear();
      frag.push_back(r);
      frag_start = line_off;
      frag_unmapped = 0;
      force_python = false;
      have_frag = true;
      g.first_record = false;
      continue;
    }
    if (r.ok && (r.flag & 0x4)) {  // unmapped records after the first are counted and skipped (find_circ.py:1466-1467)
      frag_unmapped++;
      continue;
    }
    const Rec& p = frag[0];
    if (r.ok && p.ok && r.qname == p.qname) {
      frag.push_back(r);
      continue;
    }
    if (!r.ok || !p.ok) {
      // unparsable line: make the whole neighbourhood complex and let python report it
      force_python = true;
      if (!r.ok && p.ok) {
        frag.push_back(r);
        continue;
      }
    }
    // a new fragment starts at this line: the previous one is complete
    if (!finish(line_off)) {
      // output full: stop before the fragment that did not fit
      return frag_start;
    }
    consumed = line_off;
    frag.clear();
    frag.push_back(r);
    frag_start = line_off;
    frag_unmapped = 0;
    force_python = false;
  }
  if (final && have_frag && pos >= nbytes) {
    if (!finish(nbytes)) return frag_start;
    consumed = nbytes;
    have_frag = false;
  }
  return consumed;
}

// ---------------------------------------------------------------- spliced reads: keep the text, format it at the end
// The reads file (find_circ.py:1442-1447) names the junction(s) a read supports, and junction names are only final after
// the aggregation.  For the rows of the native ingest (one junction, no flags) the host keeps name, sequence and
// qualities of every spliced read in one compact blob per batch (fc_text_gather) and formats the FASTQ records in one
// pass when the names are known (fc_fastq_format).

// ---- the evidence rules of record_hits (find_circ.py:1276-1439) for a whole batch of the native ingest -------------
// With at most two spans per fragment (back-splices first) every rule is a comparison between the fragment's columns and
// the scan's answers for its rows: one pass over the m fragment records.  pipeline.Run._record_hits stays the reading of
// the reference for everything else (three spans and more, --all-hits); the two are pinned to the same goldens.
// Outputs (arrays sized by the caller): the four hit counters; per fragment the flag word, a class byte and the junction
// keys (chrom id, start, end, minus, kind) of its spans and of its (last) back-splice; the compact list of per-junction
// evidence events (<= 2 m); the compact list of reads to write (<= 2 m: one entry per mate of a fragment with a junction).
extern "C" int fc_ingest_evidence(const fc_evidence_in* in, fc_evidence_out* o) {
  if (!in || !o || in->n < 0 || in->m < 0) return FC_E_ARG;
  const int64_t m = in->m;
  const uint32_t* B = in->bit;
  enum { UNRES_BACK = 0, CLOSURE, UNRES_LIN, OUT_SPLICE, IN_SPLICE, OTHER_CHROM, OUT_MATE, IN_MATE, BROKEN, MULTI };
  int64_t cs_n = 0, cn_n = 0, ls_n = 0, ln_n = 0, n_ev = 0, n_reads = 0;
  bool any_hit = false;
  for (int64_t f = 0; f < m; ++f) {
    const unsigned st = in->f_state[f], kd = in->f_kind[f], ff = in->f_flags[f];
    const int64_t r0 = in->f_row0[f];
    const bool two = in->f_nsp[f] == 2;
    const bool q[2] = {(st & 1u) != 0, (st & 2u) != 0};
    const int64_t row[2] = {q[0] ? r0 : 0, q[1] ? r0 + (int64_t)(st & 1u) : 0};
    const bool c[2] = {(kd & 1u) != 0, (kd & 2u) != 0};
    const bool l[2] = {!c[0], two && !c[1]};
    bool h[2];
    int64_t key[2][5];
    for (int j = 0; j < 2; ++j) {
      const fc_hit& hit = in->hits[row[j]];
      h[j] = q[j] && (hit.w2 & 0xFFFFu) != 0u;
      key[j][0] = in->chrom[row[j]];
      key[j][1] = hit.start;
      key[j][2] = hit.end;
      key[j][3] = hit.w3 & 1u;
      key[j][4] = l[j] ? 1 : 0;
      cs_n += c[j] && h[j];
      cn_n += c[j] && q[j] && !h[j];
      ls_n += l[j] && h[j];
      ln_n += l[j] && q[j] && !h[j];
    }
    any_hit = any_hit || h[0] || h[1];
    const bool ch[2] = {c[0] && h[0], c[1] && h[1]};
    const bool differ = memcmp(key[0], key[1], sizeof(key[0])) != 0;
    const bool multi = ch[0] && ch[1] && differ;  // two different back-splices (find_circ.py:1319-1329)
    const bool single = (ch[0] || ch[1]) && !multi;
    const int64_t* ck = ch[1] ? key[1] : key[0];  // the (last) back-splice of the fragment
    const int64_t cs = ck[1], ce = ck[2];
    uint32_t W = 0;
    uint8_t cls = 0;
    if ((c[0] && q[0] && !h[0]) || (c[1] && q[1] && !h[1])) W |= B[UNRES_BACK];
    if (single && c[0] && c[1]) W |= B[CLOSURE];
    for (int j = 0; j < 2; ++j) {
      if (l[j] && q[j] && !h[j]) W |= B[UNRES_LIN];
      const bool ev = l[j] && h[j] && single;
      const bool outside = key[j][1] <= cs || key[j][2] >= ce;
      if (ev) W |= outside ? B[OUT_SPLICE] : B[IN_SPLICE];
      if (ev) cls |= (uint8_t)(FC_EV_LIN0 << j);
      if (ev && outside) cls |= (uint8_t)(FC_EV_LIN0_OUT << j);
    }
    const bool un = (ff & 1u) != 0 && single;  // FR_UNSPLICED
    const bool un_other = un && (ff & 2u) != 0;  // FR_OTHER_CHROM
    const bool un_outside = un && !un_other && ((int64_t)in->f_un_pos[f] + in->asize <= cs || (int64_t)in->f_un_aend[f] - in->asize >= ce);
    if (un_other) W |= B[OTHER_CHROM];
    if (un_outside) W |= B[OUT_MATE];
    if (un && !un_other && !un_outside) W |= B[IN_MATE];
    if (single && (ff & 4u) != 0) W |= B[BROKEN];  // FR_BROKEN
    if (multi) W = B[MULTI];
    if (un) cls |= FC_EV_UN;
    if (un_other || un_outside) cls |= FC_EV_UN_OUT;
    if (h[0]) cls |= FC_EV_HIT0;
    if (h[1]) cls |= FC_EV_HIT1;
    o->W[f] = W;
    o->cls[f] = cls;
    memcpy(o->key0 + 5 * f, key[0], sizeof(key[0]));
    memcpy(o->key1 + 5 * f, key[1], sizeof(key[1]));
    memcpy(o->ck + 5 * f, ck, sizeof(key[0]));
    // per-junction flags (find_circ.py:1325-1327, 1433-1437)
    const uint64_t name_hash = in->qname_hash[r0];
    if (single && W != 0u) {
      memcpy(o->ev_key + 5 * n_ev, ck, sizeof(key[0]));
      o->ev_hash[n_ev] = name_hash;
      o->ev_mask[n_ev++] = W;
    } else if (multi) {
      for (int j = 0; j < 2; ++j) {
        memcpy(o->ev_key + 5 * n_ev, key[j], sizeof(key[0]));
        o->ev_hash[n_ev] = name_hash;
        o->ev_mask[n_ev++] = W;
      }
    }
    // the reads of every fragment with a junction (find_circ.py:1439, 1442-1447)
    if (h[0] || h[1]) {
      const int64_t* first = h[0] ? key[0] : key[1];
      const bool with2 = h[0] && h[1] && differ;
      const int n_mates = 1 + ((ff & 8u) != 0);  // FR_TWO_MATES
      for (int mate = 0; mate < n_mates; ++mate) {
        o->r_seq[n_reads] = in->f_seq[f];
        memcpy(o->r_k0 + 5 * n_reads, first, sizeof(key[0]));
        for (int k = 0; k < 5; ++k) o->r_k1[5 * n_reads + k] = with2 ? key[1][k] : -1;
        o->r_mask[n_reads] = (int64_t)W;
        for (int k = 0; k < 3; ++k) {
          o->r_off3[3 * n_reads + k] = in->f_txt_off[6 * f + 3 * mate + k] + in->text_off;
          o->r_len3[3 * n_reads + k] = in->f_txt_len[6 * f + 3 * mate + k];
        }
        ++n_reads;
      }
    }
  }
  o->counters[0] = cs_n;
  o->counters[1] = cn_n;
  o->counters[2] = ls_n;
  o->counters[3] = ln_n;
  o->n_events = n_ev;
  o->n_reads = n_reads;
  o->any_hit = any_hit ? 1 : 0;
  return FC_OK;
}

// distinct rows of an n x width matrix of 64-bit integers (np.unique(axis=0) without the sort): inverse[i] = number of
// row i's value in order of first appearance, first[u] = the row where value u appears first; returns the number of
// distinct rows.  Exact (rows are compared in full); one open-addressing table, host code.  Used by the text writers:
// the (junctions, flags) combinations of the reads, the junctions of the evidence events.
extern "C" int64_t fc_unique_rows(const int64_t* rows, int64_t n, int32_t width, int64_t* first, int32_t* inverse) {
  if (!rows || !first || !inverse || n < 0 || width <= 0 || n >= (1ll << 31)) return FC_E_ARG;
  uint64_t cap = 1024;
  while (cap < 2ull * (uint64_t)n) cap <<= 1;
  std::vector<int32_t> table(cap, -1);
  int64_t nu = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t* r = rows + i * width;
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int32_t c = 0; c < width; ++c) h = fc_mix64(h ^ (uint64_t)r[c]) + 0x632BE59BD9B4E019ULL;
    uint64_t s = h & (cap - 1);
    for (;;) {
      const int32_t u = table[s];
      if (u < 0) {
        table[s] = (int32_t)nu;
        first[nu] = i;
        inverse[i] = (int32_t)nu++;
        break;
      }
      if (memcmp(rows + first[u] * width, r, sizeof(int64_t) * (size_t)width) == 0) {
        inverse[i] = u;
        break;
      }
      s = (s + 1) & (cap - 1);
    }
  }
  return nu;
}

// copies n x 3 substrings of `buf` back to back into `out` (off/len are n x 3, row major; len < 0 = field absent);
// returns the number of bytes written
extern "C" int64_t fc_text_gather(const char* buf, int64_t n, const int64_t* off, const int32_t* len, char* out) {
  if (!buf || !off || !len || !out || n < 0) return FC_E_ARG;
  int64_t w = 0;
  for (int64_t k = 0; k < 3 * n; ++k) {
    if (len[k] > 0) {
      memcpy(out + w, buf + off[k], (size_t)len[k]);
      w += len[k];
    }
  }
  return w;
}

// FASTQ records of n reads whose (name, sequence, qualities) lie back to back in `blob` (lengths in len, n x 3; qualities
// absent = the text "None", as python prints a missing pysam field): "@<name> <tail>\n<seq>\n+<name> <tail>\n<qual>\n".
// tail ("<junction names> <flags>") = names[name_off[name_idx[i]] ...].  rec_off[0..n] receives the offset of every record in `out`.
// Returns the bytes written, or the bytes needed (> out_cap) when `out` is too small.
extern "C" int64_t fc_fastq_format(const char* blob, int64_t n, const int32_t* len, const int32_t* name_idx, const char* names,
                                   const int64_t* name_off, const int32_t* name_len, char* out, int64_t out_cap,
                                   int64_t* rec_off) {
  if (!blob || !len || !name_idx || !names || !name_off || !name_len || !rec_off || n < 0) return FC_E_ARG;
  int64_t need = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t q = len[3 * i] > 0 ? len[3 * i] : 0, s = len[3 * i + 1] > 0 ? len[3 * i + 1] : 0;
    const int64_t u = len[3 * i + 2] >= 0 ? len[3 * i + 2] : 4;
    need += 2 * (q + name_len[name_idx[i]] + 2) + 2 + s + 1 + u + 1;
  }
  if (need > out_cap || !out) return need;
  int64_t r = 0, w = 0;
  for (int64_t i = 0; i < n; ++i) {
    rec_off[i] = w;
    const int32_t q = len[3 * i] > 0 ? len[3 * i] : 0, s = len[3 * i + 1] > 0 ? len[3 * i + 1] : 0, u = len[3 * i + 2];
    const char* qp = blob + r;
    const char* sp = qp + q;
    const char* up = sp + s;
    const char* jn = names + name_off[name_idx[i]];
    const int32_t jl = name_len[name_idx[i]];
    for (int pass = 0; pass < 2; ++pass) {
      out[w++] = pass ? '+' : '@';
      memcpy(out + w, qp, (size_t)q);
      w += q;
      out[w++] = ' ';
      memcpy(out + w, jn, (size_t)jl);
      w += jl;
      out[w++] = '\n';
      if (pass == 0) {
        memcpy(out + w, sp, (size_t)s);
        w += s;
        out[w++] = '\n';
      }
    }
    if (u >= 0) {
      memcpy(out + w, up, (size_t)u);
      w += u;
      r += u;
    } else {
      memcpy(out + w, "None", 4);
      w += 4;
    }
    out[w++] = '\n';
    r += q + s;
  }
  rec_off[n] = w;
  return w;
}
