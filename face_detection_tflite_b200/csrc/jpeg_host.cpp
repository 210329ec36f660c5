// JPEG marker parser + Huffman entropy decoder (see jpeg_host.h).  ITU-T T.81 baseline / extended sequential (SOF0 / SOF1)
// and progressive (SOF2) Huffman streams, 8-bit samples, restart intervals, interleaved and non-interleaved scans.
// The output is what libjpeg keeps in its coefficient buffer: quantised coefficients per 8x8 block in natural order.
#include "jpeg_host.h"

#include <algorithm>
#include <cstring>

namespace fdt {
namespace {

const uint8_t kZigzag[64 + 16] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                   63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};   // overrun guard like libjpeg's

struct Huff {
  bool set = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  uint16_t look[512];        // 9-bit lookahead: (length << 8) | symbol, 0 = longer code
  int32_t maxcode[18];       // largest code of each length (-1: none), [17] = sentinel
  int32_t valoff[17];
  bool build() {
    int n = 0;
    for (int l = 1; l <= 16; ++l) n += bits[l];
    if (n > 256) return false;
    std::memset(look, 0, sizeof look);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valoff[l] = k - code;
      if (bits[l]) {
        for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
          if (code >= (1 << l)) return false;
          if (l <= 9) {
            const int lo = code << (9 - l), cnt = 1 << (9 - l);
            for (int j = 0; j < cnt; ++j) look[lo + j] = (uint16_t)((l << 8) | vals[k]);
          }
        }
        maxcode[l] = code - 1;
      } else {
        maxcode[l] = -1;
      }
      code <<= 1;
    }
    maxcode[17] = 0x7FFFFFFF;
    set = true;
    return true;
  }
};

struct Bits {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int n = 0;
  bool at_marker = false;      // p points at the 0xFF of a marker; the decoder sees zero bits from here on
  void fill() {
    while (n <= 56) {
      uint64_t b = 0;
      if (!at_marker) {
        if (p >= end) {
          at_marker = true;
        } else if (*p == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) { b = 0xFF; p += 2; }
          else at_marker = true;
        } else {
          b = *p++;
        }
      }
      acc |= b << (56 - n);
      n += 8;
    }
  }
  inline uint32_t peek(int k) { if (n < k) fill(); return (uint32_t)(acc >> (64 - k)); }
  inline void skip(int k) { acc <<= k; n -= k; }
  inline int get(int k) { if (k == 0) return 0; uint32_t v = peek(k); skip(k); return (int)v; }
  inline int bit() { return get(1); }
  void reset() { acc = 0; n = 0; }
};

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

inline int huff_decode(Bits& br, const Huff& h) {
  const uint32_t look = br.peek(16);
  const uint16_t e = h.look[look >> 7];
  if (e) { br.skip(e >> 8); return e & 0xFF; }
  int l = 10;
  int32_t code = (int32_t)(look >> 6);
  while (l <= 16 && code > h.maxcode[l]) { ++l; code = (int32_t)(look >> (16 - l)); }
  if (l > 16) { br.skip(16); return 0; }              // corrupt data: libjpeg warns and returns 0
  br.skip(l);
  return h.vals[(code + h.valoff[l]) & 0xFF];
}

struct Decoder {
  const uint8_t* data;
  size_t len;
  JpegImage* img;
  std::string* err;
  Huff dc[4], ac[4];
  int restart_interval = 0;
  bool saw_sof = false, saw_jfif = false, saw_adobe = false;
  int adobe_transform = 0;
  std::vector<uint8_t> coef_bits_dummy;

  JpegStatus bad(const char* m) { if (err) *err = m; return kJpegBad; }
  JpegStatus unsupported(const char* m) { if (err) *err = m; return kJpegUnsupported; }

  static uint32_t rd16(const uint8_t* p, bool le) { return le ? (uint32_t)(p[0] | (p[1] << 8)) : (uint32_t)((p[0] << 8) | p[1]); }
  static uint32_t rd32(const uint8_t* p, bool le) {
    return le ? (uint32_t)(p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24)) : (uint32_t)(((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]);
  }

  void parse_exif(const uint8_t* p, size_t n) {
    if (n < 14 || std::memcmp(p, "Exif\0\0", 6) != 0) return;
    const uint8_t* t = p + 6;
    const size_t tn = n - 6;
    bool le;
    if (t[0] == 'I' && t[1] == 'I') le = true; else if (t[0] == 'M' && t[1] == 'M') le = false; else return;
    if (rd16(t + 2, le) != 42) return;
    const uint32_t ifd = rd32(t + 4, le);
    if (ifd > tn || tn - ifd < 2) return;
    const uint32_t cnt = rd16(t + ifd, le);
    for (uint32_t i = 0; i < cnt; ++i) {
      const size_t o = (size_t)ifd + 2 + 12 * (size_t)i;
      if (o + 12 > tn) return;
      if (rd16(t + o, le) == 0x0112) {
        const uint32_t v = rd16(t + o + 8, le);
        if (v >= 1 && v <= 8) img->orientation = (int)v;
        return;
      }
    }
  }

  // ---- one scan ---------------------------------------------------------------------------------------------------
  struct Scan { int ncomp; int ci[3]; int td[3], ta[3]; int Ss, Se, Ah, Al; };

  void restart(Bits& br, int* pred, int* eobrun) {
    br.reset();
    // the marker should be right here; be lenient and look for the next RSTn
    const uint8_t* q = br.p;
    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) {
      if (q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF) { q = nullptr; break; }    // some other marker: no more restarts in this scan
      ++q;
    }
    if (q && q + 1 < br.end) { br.p = q + 2; br.at_marker = false; }
    pred[0] = pred[1] = pred[2] = 0;
    *eobrun = 0;
  }

  // baseline / sequential block
  inline void block_seq(Bits& br, int16_t* c, int& pred, const Huff& hd, const Huff& ha) {
    int s = huff_decode(br, hd);
    if (s) { s &= 15; pred += extend(br.get(s), s); }
    c[0] = (int16_t)pred;
    for (int k = 1; k < 64;) {
      const int rs = huff_decode(br, ha);
      const int r = rs >> 4, sz = rs & 15;
      if (sz) {
        k += r;
        c[kZigzag[k]] = (int16_t)extend(br.get(sz), sz);
        ++k;
      } else {
        if (r != 15) break;
        k += 16;
      }
    }
  }

  inline void block_dc_first(Bits& br, int16_t* c, int& pred, const Huff& hd, int Al) {
    int s = huff_decode(br, hd);
    if (s) { s &= 15; pred += extend(br.get(s), s); }
    c[0] = (int16_t)(pred * (1 << Al));
  }
  inline void block_dc_refine(Bits& br, int16_t* c, int Al) {
    if (br.bit()) c[0] |= (int16_t)(1 << Al);
  }
  inline void block_ac_first(Bits& br, int16_t* c, const Huff& ha, int Ss, int Se, int Al, int& eobrun) {
    if (eobrun > 0) { --eobrun; return; }
    for (int k = Ss; k <= Se; ++k) {
      const int rs = huff_decode(br, ha);
      const int r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        c[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << Al));
      } else {
        if (r == 15) { k += 15; }
        else {
          eobrun = 1 << r;
          if (r) eobrun += br.get(r);
          --eobrun;
          break;
        }
      }
    }
  }
  inline void block_ac_refine(Bits& br, int16_t* c, const Huff& ha, int Ss, int Se, int Al, int& eobrun) {
    const int p1 = 1 << Al, m1 = -(1 << Al);
    int k = Ss;
    if (eobrun == 0) {
      for (; k <= Se; ++k) {
        const int rs = huff_decode(br, ha);
        int r = rs >> 4, s = rs & 15;
        if (s) {
          s = br.bit() ? p1 : m1;
        } else if (r != 15) {
          eobrun = 1 << r;
          if (r) eobrun += br.get(r);
          break;
        }
        do {
          int16_t* cf = c + kZigzag[k];
          if (*cf != 0) {
            if (br.bit()) {
              if ((*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
            }
          } else {
            if (--r < 0) break;
          }
          ++k;
        } while (k <= Se);
        if (s) c[kZigzag[k]] = (int16_t)s;
      }
    }
    if (eobrun > 0) {
      for (; k <= Se; ++k) {
        int16_t* cf = c + kZigzag[k];
        if (*cf != 0) {
          if (br.bit()) {
            if ((*cf & p1) == 0) *cf = (int16_t)(*cf + (*cf >= 0 ? p1 : m1));
          }
        }
      }
      --eobrun;
    }
  }

  // returns the position right after the entropy-coded segment (at the next marker's 0xFF)
  JpegStatus decode_scan(const Scan& sc, size_t pos, size_t* next) {
    Bits br;
    br.p = data + pos;
    br.end = data + len;
    int pred[3] = {0, 0, 0};
    int eobrun = 0;
    const bool prog = img->progressive;
    for (int i = 0; i < sc.ncomp; ++i) {
      const bool need_dc = !prog || sc.Ss == 0, need_ac = !prog || sc.Ss > 0;
      if (need_dc && !(prog && sc.Ah) && !dc[sc.td[i]].set) return bad("JPEG: scan uses an undefined DC Huffman table");
      if (need_ac && !ac[sc.ta[i]].set) return bad("JPEG: scan uses an undefined AC Huffman table");
    }
    auto do_block = [&](int i, int16_t* c) {
      JpegComp& cp = img->comp[sc.ci[i]];
      (void)cp;
      if (!prog) block_seq(br, c, pred[i], dc[sc.td[i]], ac[sc.ta[i]]);
      else if (sc.Ss == 0) { if (sc.Ah == 0) block_dc_first(br, c, pred[i], dc[sc.td[i]], sc.Al); else block_dc_refine(br, c, sc.Al); }
      else { if (sc.Ah == 0) block_ac_first(br, c, ac[sc.ta[i]], sc.Ss, sc.Se, sc.Al, eobrun); else block_ac_refine(br, c, ac[sc.ta[i]], sc.Ss, sc.Se, sc.Al, eobrun); }
    };
    int mcu = 0;
    if (sc.ncomp == 1) {
      // non-interleaved: the component's own blocks in raster order (not padded to MCUs)
      JpegComp& cp = img->comp[sc.ci[0]];
      const int nbx = (cp.dw + 7) / 8, nby = (cp.dh + 7) / 8;
      for (int by = 0; by < nby; ++by)
        for (int bx = 0; bx < nbx; ++bx, ++mcu) {
          if (restart_interval && mcu && mcu % restart_interval == 0) restart(br, pred, &eobrun);
          do_block(0, cp.coef.data() + ((size_t)by * cp.bw + bx) * 64);
        }
    } else {
      for (int my = 0; my < img->mcuy; ++my)
        for (int mx = 0; mx < img->mcux; ++mx, ++mcu) {
          if (restart_interval && mcu && mcu % restart_interval == 0) restart(br, pred, &eobrun);
          for (int i = 0; i < sc.ncomp; ++i) {
            JpegComp& cp = img->comp[sc.ci[i]];
            for (int v = 0; v < cp.v; ++v)
              for (int h = 0; h < cp.h; ++h)
                do_block(i, cp.coef.data() + ((size_t)(my * cp.v + v) * cp.bw + (mx * cp.h + h)) * 64);
          }
        }
    }
    // position of the next marker
    const uint8_t* q = br.at_marker ? br.p : br.p;
    // bytes already pulled into the accumulator lie before br.p only when no marker was hit; scan forward to a real marker
    while (q + 1 < br.end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) ++q;
    *next = (size_t)(q - data);
    return kJpegOk;
  }

  JpegStatus run() {
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return bad("not a JPEG stream (no SOI marker)");
    size_t pos = 2;
    bool done = false, any_scan = false;
    while (!done) {
      // next marker
      while (pos < len && data[pos] != 0xFF) ++pos;
      while (pos < len && data[pos] == 0xFF) ++pos;
      if (pos >= len) break;
      const int mk = data[pos++];
      if (mk == 0xD9) { done = true; break; }
      if (mk == 0x01 || (mk >= 0xD0 && mk <= 0xD7)) continue;
      if (pos + 2 > len) break;
      const size_t L = ((size_t)data[pos] << 8) | data[pos + 1];
      if (L < 2 || pos + L > len) { if (any_scan) break; return bad("JPEG: truncated marker segment"); }
      const uint8_t* seg = data + pos + 2;
      const size_t sl = L - 2;
      switch (mk) {
        case 0xE0: if (sl >= 5 && std::memcmp(seg, "JFIF\0", 5) == 0) saw_jfif = true; break;
        case 0xE1: parse_exif(seg, sl); break;
        case 0xEE: if (sl >= 12 && std::memcmp(seg, "Adobe", 5) == 0) { saw_adobe = true; adobe_transform = seg[11]; } break;
        case 0xDB: {
          size_t o = 0;
          while (o < sl) {
            const int pq = seg[o] >> 4, tq = seg[o] & 15;
            ++o;
            if (tq > 3 || pq > 1) return bad("JPEG: bad quantisation table header");
            const size_t need = pq ? 128 : 64;
            if (o + need > sl) return bad("JPEG: truncated quantisation table");
            for (int i = 0; i < 64; ++i) {
              const uint32_t v = pq ? (uint32_t)((seg[o + 2 * i] << 8) | seg[o + 2 * i + 1]) : seg[o + i];
              img->qt[tq][kZigzag[i]] = (uint16_t)v;
            }
            img->qt_set[tq] = true;
            o += need;
          }
          break;
        }
        case 0xC4: {
          size_t o = 0;
          while (o < sl) {
            if (o + 17 > sl) return bad("JPEG: truncated Huffman table");
            const int tc = seg[o] >> 4, th = seg[o] & 15;
            if (tc > 1 || th > 3) return bad("JPEG: bad Huffman table header");
            Huff& h = tc ? ac[th] : dc[th];
            int n = 0;
            h.bits[0] = 0;
            for (int i = 1; i <= 16; ++i) { h.bits[i] = seg[o + i]; n += seg[o + i]; }
            o += 17;
            if (n > 256 || o + (size_t)n > sl) return bad("JPEG: bad Huffman table");
            std::memset(h.vals, 0, sizeof h.vals);
            std::memcpy(h.vals, seg + o, (size_t)n);
            o += (size_t)n;
            if (!h.build()) return bad("JPEG: inconsistent Huffman table");
          }
          break;
        }
        case 0xDD: if (sl >= 2) restart_interval = (seg[0] << 8) | seg[1]; break;
        case 0xC0: case 0xC1: case 0xC2: {
          if (saw_sof) return bad("JPEG: more than one frame header");
          if (sl < 6) return bad("JPEG: truncated frame header");
          if (seg[0] != 8) return unsupported("JPEG: only 8-bit samples are supported");
          img->progressive = mk == 0xC2;
          img->height = (seg[1] << 8) | seg[2];
          img->width = (seg[3] << 8) | seg[4];
          img->ncomp = seg[5];
          if (img->ncomp != 1 && img->ncomp != 3) return unsupported("JPEG: only 1- and 3-component images are supported");
          if (img->width <= 0 || img->height <= 0) return bad("JPEG: empty image (DNL-defined heights are not supported)");
          if ((long long)img->width * img->height > (1LL << 26)) return unsupported("JPEG: image larger than 64 MP");
          if (sl < 6 + 3 * (size_t)img->ncomp) return bad("JPEG: truncated frame header");
          img->hmax = img->vmax = 1;
          for (int i = 0; i < img->ncomp; ++i) {
            JpegComp& c = img->comp[i];
            c.id = seg[6 + 3 * i]; c.h = seg[7 + 3 * i] >> 4; c.v = seg[7 + 3 * i] & 15; c.tq = seg[8 + 3 * i];
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return bad("JPEG: bad component parameters");
            img->hmax = std::max(img->hmax, c.h); img->vmax = std::max(img->vmax, c.v);
          }
          if (img->ncomp == 1) { img->comp[0].h = img->comp[0].v = 1; img->hmax = img->vmax = 1; }   // a single component is never subsampled
          img->mcux = (img->width + 8 * img->hmax - 1) / (8 * img->hmax);
          img->mcuy = (img->height + 8 * img->vmax - 1) / (8 * img->vmax);
          for (int i = 0; i < img->ncomp; ++i) {
            JpegComp& c = img->comp[i];
            if (img->ncomp == 3 && !((c.h == img->hmax || c.h * 2 == img->hmax) && (c.v == img->vmax || c.v * 2 == img->vmax)))
              return unsupported("JPEG: only 1x and 2x chroma subsampling is supported");
            c.dw = (img->width * c.h + img->hmax - 1) / img->hmax;
            c.dh = (img->height * c.v + img->vmax - 1) / img->vmax;
            c.bw = img->mcux * c.h;
            c.bh = img->mcuy * c.v;
            c.coef.assign((size_t)c.bw * c.bh * 64, 0);
          }
          if (img->ncomp == 3 && (img->comp[0].h != img->hmax || img->comp[0].v != img->vmax || img->comp[1].h != img->comp[2].h || img->comp[1].v != img->comp[2].v))
            return unsupported("JPEG: luma must be the full-resolution component and both chroma components must share a sampling factor");
          saw_sof = true;
          break;
        }
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          return unsupported("JPEG: lossless, hierarchical and arithmetic-coded streams are not supported");
        case 0xDA: {
          if (!saw_sof) return bad("JPEG: scan before the frame header");
          if (sl < 1) return bad("JPEG: truncated scan header");
          Scan sc;
          sc.ncomp = seg[0];
          if (sc.ncomp < 1 || sc.ncomp > img->ncomp || sl < 1 + 2 * (size_t)sc.ncomp + 3) return bad("JPEG: bad scan header");
          for (int i = 0; i < sc.ncomp; ++i) {
            const int id = seg[1 + 2 * i];
            int ci = -1;
            for (int k = 0; k < img->ncomp; ++k) if (img->comp[k].id == id) ci = k;
            if (ci < 0) return bad("JPEG: scan names an unknown component");
            sc.ci[i] = ci; sc.td[i] = seg[2 + 2 * i] >> 4; sc.ta[i] = seg[2 + 2 * i] & 15;
            if (sc.td[i] > 3 || sc.ta[i] > 3) return bad("JPEG: bad Huffman table selector");
          }
          const uint8_t* t = seg + 1 + 2 * sc.ncomp;
          sc.Ss = t[0]; sc.Se = t[1]; sc.Ah = t[2] >> 4; sc.Al = t[2] & 15;
          if (img->progressive) {
            if (sc.Ss > sc.Se || sc.Se > 63 || sc.Al > 13 || (sc.Ss == 0 && sc.Se != 0) || (sc.Ss > 0 && sc.ncomp != 1)) return bad("JPEG: bad progressive scan parameters");
          } else {
            sc.Ss = 0; sc.Se = 63; sc.Ah = sc.Al = 0;
          }
          size_t next = 0;
          JpegStatus st = decode_scan(sc, pos + L, &next);
          if (st != kJpegOk) return st;
          any_scan = true;
          pos = next;
          continue;
        }
        default: break;
      }
      pos += L;
    }
    if (!saw_sof || !any_scan) return bad("JPEG: no image data");
    for (int i = 0; i < img->ncomp; ++i)
      if (!img->qt_set[img->comp[i].tq]) return bad("JPEG: a component uses an undefined quantisation table");
    if (img->ncomp == 3) {
      // libjpeg's colour-space guess (jdapimin.c default_decompress_parms): everything that is not YCbCr is refused
      bool ycc = true;
      if (saw_jfif) ycc = true;
      else if (saw_adobe) ycc = adobe_transform != 0;
      else if (img->comp[0].id == 'R' && img->comp[1].id == 'G' && img->comp[2].id == 'B') ycc = false;
      if (!ycc) return unsupported("JPEG: RGB-coded (non-YCbCr) streams are not supported");
    }
    return kJpegOk;
  }
};

}  // namespace

JpegStatus jpeg_decode_coefficients(const uint8_t* data, size_t n, JpegImage* img, std::string* err) {
  if (!data || !img) { if (err) *err = "null JPEG buffer"; return kJpegBad; }
  Decoder d;
  d.data = data; d.len = n; d.img = img; d.err = err;
  *img = JpegImage();
  std::memset(img->qt, 0, sizeof img->qt);
  return d.run();
}

}  // namespace fdt
