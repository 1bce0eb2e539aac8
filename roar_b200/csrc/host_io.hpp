// Host-side I/O of the extraction run, native and multi-threaded (no GIL): RIFF/WAVE decoding straight into
// the caller's (pinned) staging buffers, and the `.pt` cache files the reference writes with torch.save.
//
// Replaces, on the host side of the path:
//   * AudioSegment.from_file for wav input (asr/parts/preprocessing/segment.py:156-278: soundfile read as
//     float32, `offset` / `duration` in seconds, integer PCM scaled by 2^-(bits-1), :140-153) as reached through
//     WaveformFeaturizer.process (asr/parts/preprocessing/features.py:137-170) from TTSDataset.__getitem__
//     (tts/data/dataset.py:613-621);
//   * torch.save(tensor, path) of the cached sup data (dataset.py:656-657, 704-708, 752-753): a stored
//     (uncompressed) ZIP archive holding `<root>/data.pkl` (pickle protocol 2 calling
//     torch._utils._rebuild_tensor_v2 on a persistent-id FloatStorage), `<root>/byteorder`, `<root>/data/0` (raw
//     little-endian float32, 64-byte aligned) and `<root>/version` -- what torch.load reads back.
// Pure C++17 (no CUDA), so the CPU tests exercise it without a GPU.
#pragma once
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include "../../include/roar_sup.h"

namespace roar_io {

// ---------------------------------------------------------------------------------------------- helpers
inline uint32_t rd_u32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd_u16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// run fn(i) for i in [0, n) on up to n_threads threads (dynamic work distribution)
template <class F> inline void parallel_for(int n, int n_threads, F fn) {
  if (n <= 0) return;
  if (n_threads > n) n_threads = n;
  if (n_threads <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  th.reserve(n_threads);
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i); } });
  for (auto& t : th) t.join();
}

inline bool pread_all(int fd, void* buf, size_t n, int64_t off) {
  unsigned char* p = (unsigned char*)buf;
  while (n > 0) {
    const ssize_t r = ::pread(fd, p, n, off);
    if (r < 0) { if (errno == EINTR) continue; return false; }
    if (r == 0) return false;
    p += r; n -= (size_t)r; off += r;
  }
  return true;
}
inline bool write_all(int fd, const void* buf, size_t n) {
  const unsigned char* p = (const unsigned char*)buf;
  while (n > 0) {
    const ssize_t r = ::write(fd, p, n);
    if (r < 0) { if (errno == EINTR) continue; return false; }
    p += r; n -= (size_t)r;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------- WAV
// -> 0 ok; otherwise info->sample_rate = -1 and a reason in *err
inline int wav_probe(const char* path, roar_wav_info* info, std::string* err) {
  memset(info, 0, sizeof(*info));
  info->sample_rate = -1;
  const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) { if (err) *err = std::string(path) + ": " + strerror(errno); return -1; }
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); if (err) *err = std::string(path) + ": fstat failed"; return -1; }
  const int64_t fsize = st.st_size;
  unsigned char hdr[12];
  if (fsize < 12 || !pread_all(fd, hdr, 12, 0) || memcmp(hdr, "RIFF", 4) != 0 || memcmp(hdr + 8, "WAVE", 4) != 0) {
    ::close(fd);
    if (err) *err = std::string(path) + ": not a RIFF/WAVE file (only wav input is decoded natively)";
    return -1;
  }
  int64_t pos = 12;
  bool have_fmt = false;
  int fmt_tag = 0, channels = 0, bits = 0, block = 0, rate = 0;
  while (pos + 8 <= fsize) {
    unsigned char ch[8];
    if (!pread_all(fd, ch, 8, pos)) break;
    const uint32_t sz = rd_u32(ch + 4);
    if (memcmp(ch, "fmt ", 4) == 0) {
      unsigned char f[40];
      const size_t n = sz < sizeof(f) ? sz : sizeof(f);
      if (n < 16 || !pread_all(fd, f, n, pos + 8)) break;
      fmt_tag = rd_u16(f); channels = rd_u16(f + 2); rate = (int)rd_u32(f + 4); block = rd_u16(f + 12); bits = rd_u16(f + 14);
      if (fmt_tag == 0xFFFE && n >= 26) fmt_tag = rd_u16(f + 24);      // WAVE_FORMAT_EXTENSIBLE: sub-format GUID
      have_fmt = true;
    } else if (memcmp(ch, "data", 4) == 0) {
      if (!have_fmt) break;
      int64_t bytes = sz;
      if (sz == 0 || sz == 0xFFFFFFFFu || pos + 8 + bytes > fsize) bytes = fsize - (pos + 8);   // streamed / truncated
      ::close(fd);
      const bool pcm = fmt_tag == 1 && (bits == 8 || bits == 16 || bits == 24 || bits == 32);
      const bool flt = fmt_tag == 3 && (bits == 32 || bits == 64);
      if (!(pcm || flt) || channels < 1 || block != channels * bits / 8 || rate <= 0) {
        if (err) *err = std::string(path) + ": unsupported wav encoding (format " + std::to_string(fmt_tag) + ", " + std::to_string(bits) + " bit)";
        return -1;
      }
      info->sample_rate = rate; info->channels = channels; info->bits = bits; info->format = fmt_tag;
      info->n_frames = bytes / block; info->data_offset = pos + 8;
      return 0;
    }
    pos += 8 + (int64_t)sz + (sz & 1);
  }
  ::close(fd);
  if (err) *err = std::string(path) + ": malformed wav (no fmt/data chunk)";
  return -1;
}

// one sample of any supported encoding -> float32 the way libsndfile's float read does
inline float wav_sample(const unsigned char* p, int format, int bits) {
  if (format == 3) {
    if (bits == 32) { float f; memcpy(&f, p, 4); return f; }
    double d; memcpy(&d, p, 8); return (float)d;
  }
  switch (bits) {
    case 8: return ((float)p[0] - 128.0f) * (1.0f / 128.0f);
    case 16: return (float)(int16_t)rd_u16(p) * (1.0f / 32768.0f);
    case 24: { int32_t v = (int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24) >> 8; return (float)v * (1.0f / 8388608.0f); }
    default: return (float)((double)(int32_t)rd_u32(p) * (1.0 / 2147483648.0));
  }
}

// frames [first, first + count) of one file -> float32 mono at dst.  channel: -1 average, else that channel.
inline int wav_read_f32(const char* path, const roar_wav_info& w, int64_t first, int64_t count, int channel, float* dst, std::string* err) {
  if (w.sample_rate <= 0 || first < 0 || count < 0 || first + count > w.n_frames || channel >= w.channels) {
    if (err) *err = std::string(path) + ": bad read range / channel";
    return -1;
  }
  if (count == 0) return 0;
  const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) { if (err) *err = std::string(path) + ": " + strerror(errno); return -1; }
  const int bps = w.bits / 8, block = bps * w.channels;
  int rc = 0;
  if (w.format == 3 && w.bits == 32 && w.channels == 1) {          // float32 mono: read in place
    if (!pread_all(fd, dst, (size_t)count * 4, w.data_offset + first * 4)) rc = -1;
  } else {
    const int64_t chunk = 1 << 16;
    std::vector<unsigned char> buf((size_t)chunk * block);
    for (int64_t done = 0; done < count && rc == 0; done += chunk) {
      const int64_t n = count - done < chunk ? count - done : chunk;
      if (!pread_all(fd, buf.data(), (size_t)n * block, w.data_offset + (first + done) * block)) { rc = -1; break; }
      for (int64_t i = 0; i < n; ++i) {
        const unsigned char* fr = buf.data() + i * block;
        if (w.channels == 1) dst[done + i] = wav_sample(fr, w.format, w.bits);
        else if (channel >= 0) dst[done + i] = wav_sample(fr + channel * bps, w.format, w.bits);
        else {          // np.mean over the channel axis of the float32 samples
          float acc = 0.f;
          for (int c = 0; c < w.channels; ++c) acc += wav_sample(fr + c * bps, w.format, w.bits);
          dst[done + i] = acc / (float)w.channels;
        }
      }
    }
  }
  ::close(fd);
  if (rc != 0 && err) *err = std::string(path) + ": short read";
  return rc;
}

// 16-bit mono PCM, raw: the integer -> float step happens on the GPU (roar_sup_pcm16_to_f32)
inline int wav_read_pcm16(const char* path, const roar_wav_info& w, int64_t first, int64_t count, int16_t* dst, std::string* err) {
  if (w.sample_rate <= 0 || w.format != 1 || w.bits != 16 || w.channels != 1 || first < 0 || count < 0 || first + count > w.n_frames) {
    if (err) *err = std::string(path) + ": not 16-bit mono PCM / bad range";
    return -1;
  }
  if (count == 0) return 0;
  const int fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) { if (err) *err = std::string(path) + ": " + strerror(errno); return -1; }
  const bool ok = pread_all(fd, dst, (size_t)count * 2, w.data_offset + first * 2);
  ::close(fd);
  if (!ok && err) *err = std::string(path) + ": short read";
  return ok ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------- CRC-32 (zip)
struct Crc32 {
  uint32_t t[8][256];
  Crc32() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
  uint32_t operator()(const void* data, size_t n, uint32_t crc = 0) const {
    const unsigned char* p = (const unsigned char*)data;
    crc = ~crc;
    while (n >= 8) {
      uint32_t a, b;
      memcpy(&a, p, 4); memcpy(&b, p + 4, 4);
      a ^= crc;
      crc = t[7][a & 0xff] ^ t[6][(a >> 8) & 0xff] ^ t[5][(a >> 16) & 0xff] ^ t[4][a >> 24] ^
            t[3][b & 0xff] ^ t[2][(b >> 8) & 0xff] ^ t[1][(b >> 16) & 0xff] ^ t[0][b >> 24];
      p += 8; n -= 8;
    }
    while (n--) crc = t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return ~crc;
  }
};
inline const Crc32& crc32_tab() { static const Crc32 c; return c; }

// ---------------------------------------------------------------------------------------------- .pt writer
inline void put_u16(std::string& s, uint32_t v) { s.push_back((char)(v & 0xff)); s.push_back((char)((v >> 8) & 0xff)); }
inline void put_u32(std::string& s, uint32_t v) { put_u16(s, v & 0xffff); put_u16(s, v >> 16); }
inline void pkl_int(std::string& s, int64_t v) {       // what pickle protocol 2 emits for a non-negative int
  if (v < 256) { s.push_back('K'); s.push_back((char)v); }
  else if (v < 65536) { s.push_back('M'); put_u16(s, (uint32_t)v); }
  else { s.push_back('J'); put_u32(s, (uint32_t)v); }      // < 2^31
}
inline void pkl_tuple(std::string& s, const int64_t* v, int n) {
  for (int i = 0; i < n; ++i) pkl_int(s, v[i]);
  s.push_back(n == 1 ? '\x85' : n == 2 ? '\x86' : '\x87');
}
// data.pkl of torch.save(float32 CPU tensor of `shape`, contiguous)
inline std::string pt_pickle(const int64_t* shape, int rank) {
  int64_t numel = 1, stride[3] = {1, 1, 1};
  for (int i = rank - 1; i >= 0; --i) { stride[i] = numel; numel *= shape[i]; }
  std::string s;
  s += "\x80\x02" "ctorch._utils\n_rebuild_tensor_v2\nq";
  s.push_back('\0');
  s += "((X\x07"; s.append(3, '\0'); s += "storageq\x01" "ctorch\nFloatStorage\nq\x02" "X\x01"; s.append(3, '\0'); s += "0q\x03" "X\x03"; s.append(3, '\0'); s += "cpuq\x04";
  pkl_int(s, numel);
  s += "tq\x05Q";
  s.push_back('K'); s.push_back('\0');            // storage offset
  pkl_tuple(s, shape, rank);
  s += "q\x06";
  pkl_tuple(s, stride, rank);
  s += "q\x07\x89" "ccollections\nOrderedDict\nq\x08)Rq\ttq\nRq\x0b.";
  return s;
}

struct ZipEntry { std::string name; uint32_t crc, size, offset; };
inline void zip_local(std::string& out, std::vector<ZipEntry>& dir, const std::string& name, const void* data, uint32_t size,
                      uint32_t crc, bool align64) {
  ZipEntry e; e.name = name; e.crc = crc; e.size = size; e.offset = (uint32_t)out.size();
  // optional "FB" extra field pads the payload to a 64-byte boundary like torch's writer
  size_t extra = 0;
  if (align64) {
    const size_t start = out.size() + 30 + name.size();
    extra = (64 - (start + 4) % 64) % 64 + 4;
  }
  put_u32(out, 0x04034b50); put_u16(out, 20); put_u16(out, 0x0800); put_u16(out, 0); put_u16(out, 0); put_u16(out, 0x21);
  put_u32(out, crc); put_u32(out, size); put_u32(out, size); put_u16(out, (uint32_t)name.size()); put_u16(out, (uint32_t)extra);
  out += name;
  if (extra) { out += "FB"; put_u16(out, (uint32_t)(extra - 4)); out.append(extra - 4, 'Z'); }
  if (data) out.append((const char*)data, size);
  dir.push_back(e);
}
inline void zip_finish(std::string& out, const std::vector<ZipEntry>& dir, size_t payload_after) {
  // central directory offsets account for `payload_after` bytes written separately after `out`'s local part
  (void)payload_after;
  const uint32_t cd_off = (uint32_t)out.size();
  for (const auto& e : dir) {
    put_u32(out, 0x02014b50); put_u16(out, 20); put_u16(out, 20); put_u16(out, 0x0800); put_u16(out, 0); put_u16(out, 0); put_u16(out, 0x21);
    put_u32(out, e.crc); put_u32(out, e.size); put_u32(out, e.size); put_u16(out, (uint32_t)e.name.size()); put_u16(out, 0); put_u16(out, 0);
    put_u16(out, 0); put_u16(out, 0); put_u32(out, 0); put_u32(out, e.offset);
    out += e.name;
  }
  const uint32_t cd_size = (uint32_t)out.size() - cd_off;
  put_u32(out, 0x06054b50); put_u16(out, 0); put_u16(out, 0); put_u16(out, (uint32_t)dir.size()); put_u16(out, (uint32_t)dir.size());
  put_u32(out, cd_size); put_u32(out, cd_off); put_u16(out, 0);
}

// archive = [local headers + small entries | tensor bytes | trailing entries + central directory]: the tensor
// bytes go from the caller's buffer to the file by writev (no staging copy); temp name, rename into place
inline int pt_write_f32(const char* path, const float* data, const int64_t* shape, int rank, std::string* err) {
  if (rank < 1 || rank > 3) { if (err) *err = "rank must be 1..3"; return -1; }
  int64_t numel = 1;
  for (int i = 0; i < rank; ++i) { if (shape[i] < 0) { if (err) *err = "negative dimension"; return -1; } numel *= shape[i]; }
  if (numel >= ((int64_t)1 << 29)) { if (err) *err = std::string(path) + ": tensor too large for the native .pt writer"; return -1; }
  const uint32_t nbytes = (uint32_t)(numel * 4);
  const Crc32& crc = crc32_tab();
  const std::string root = "archive/";
  const std::string pkl = pt_pickle(shape, rank);
  std::string head, tail;
  head.reserve(512); tail.reserve(512);
  std::vector<ZipEntry> dir;
  zip_local(head, dir, root + "data.pkl", pkl.data(), (uint32_t)pkl.size(), crc(pkl.data(), pkl.size()), true);
  zip_local(head, dir, root + "byteorder", "little", 6, crc("little", 6), false);
  zip_local(head, dir, root + "data/0", nullptr, nbytes, crc(data, nbytes), true);     // payload follows by writev
  const size_t base = head.size() + nbytes;
  {
    std::string t;       // entries after the payload: offsets are relative to the whole file
    std::vector<ZipEntry> d2;
    zip_local(t, d2, root + "version", "3\n", 2, crc("3\n", 2), false);
    d2[0].offset += (uint32_t)base;
    dir.push_back(d2[0]);
    tail = t;
  }
  {
    std::string cd;
    zip_finish(cd, dir, 0);
    // zip_finish wrote the central-directory offset relative to `cd`: patch it to the file position
    const uint32_t cd_off = (uint32_t)(base + tail.size());
    const size_t eocd = cd.size() - 22;
    cd[eocd + 16] = (char)(cd_off & 0xff); cd[eocd + 17] = (char)((cd_off >> 8) & 0xff);
    cd[eocd + 18] = (char)((cd_off >> 16) & 0xff); cd[eocd + 19] = (char)((cd_off >> 24) & 0xff);
    tail += cd;
  }
  const std::string tmp = std::string(path) + ".tmp" + std::to_string((long)getpid()) + "_" + std::to_string((unsigned long)(uintptr_t)data & 0xffff);
  const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
  if (fd < 0) { if (err) *err = tmp + ": " + strerror(errno); return -1; }
  struct iovec iov[3];
  iov[0].iov_base = (void*)head.data(); iov[0].iov_len = head.size();
  iov[1].iov_base = (void*)data; iov[1].iov_len = nbytes;
  iov[2].iov_base = (void*)tail.data(); iov[2].iov_len = tail.size();
  const size_t total = head.size() + nbytes + tail.size();
  bool ok = true;
  const ssize_t w = ::writev(fd, iov, 3);
  if (w < 0) ok = false;
  else if ((size_t)w < total) {            // partial writev: finish with plain writes
    size_t done = (size_t)w;
    for (int k = 0; k < 3 && ok; ++k) {
      if (done >= iov[k].iov_len) { done -= iov[k].iov_len; continue; }
      ok = write_all(fd, (const char*)iov[k].iov_base + done, iov[k].iov_len - done);
      done = 0;
    }
  }
  ::close(fd);
  if (!ok || ::rename(tmp.c_str(), path) != 0) {
    if (err) *err = std::string(path) + ": " + strerror(errno);
    ::unlink(tmp.c_str());
    return -1;
  }
  return 0;
}

}  // namespace roar_io
