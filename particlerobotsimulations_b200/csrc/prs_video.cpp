/*
 * prs_video.cpp — video file of headless frames.
 *
 * The reference hands every VIDEO_INTERVAL-th displayed frame to cv::VideoWriter (XVID, 20 frames per second,
 * postprocess.cu:22-25, 101-118).  Neither OpenCV nor a codec exists in a headless build, so the container is written
 * directly: an AVI (RIFF) file with one uncompressed video stream — BITMAPINFOHEADER with BI_RGB, 24 bits per pixel, rows
 * bottom-up and padded to 4 bytes, one '00db' chunk per frame and an 'idx1' index — which every player and ffmpeg read.
 * Frames come in as the top-down B, G, R rows prs_render_frame produces (the layout of the reference's cv::Mat frame) and
 * are flipped on the way out.  Host code only.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "prs_cabi.h"

struct prs_video {
  FILE *fp = nullptr;
  uint32_t width = 0, height = 0, row_bytes = 0, frame_bytes = 0, frames = 0;
  uint32_t usec_per_frame = 50000, rate = 20, scale = 1;
  long movi_list_pos = 0; /* file offset of the 'LIST' word of the movi list */
  std::vector<uint32_t> offsets; /* of every frame chunk, relative to the 'movi' word */
  std::vector<unsigned char> row;
  bool failed = false;
};

namespace {
void put32(FILE *fp, uint32_t v) {
  unsigned char b[4] = {(unsigned char)v, (unsigned char)(v >> 8), (unsigned char)(v >> 16), (unsigned char)(v >> 24)};
  fwrite(b, 1, 4, fp);
}
void put16(FILE *fp, uint16_t v) {
  unsigned char b[2] = {(unsigned char)v, (unsigned char)(v >> 8)};
  fwrite(b, 1, 2, fp);
}
void tag(FILE *fp, const char *t) { fwrite(t, 1, 4, fp); }

/* header with the frame count and sizes known so far; rewritten by close() */
void write_header(prs_video *v) {
  FILE *fp = v->fp;
  const uint32_t movi_bytes = 4 + v->frames * (8 + v->frame_bytes);
  const uint32_t idx_bytes = v->frames * 16;
  const uint32_t hdrl_bytes = 4 + (8 + 56) + (8 + 4 + (8 + 56) + (8 + 40));
  const uint32_t riff_bytes = 4 + (8 + hdrl_bytes) + (8 + movi_bytes) + (8 + idx_bytes);
  fseek(fp, 0, SEEK_SET);
  tag(fp, "RIFF"); put32(fp, riff_bytes); tag(fp, "AVI ");
  tag(fp, "LIST"); put32(fp, hdrl_bytes); tag(fp, "hdrl");
  tag(fp, "avih"); put32(fp, 56);
  put32(fp, v->usec_per_frame);
  put32(fp, (uint32_t)((uint64_t)v->frame_bytes * v->rate / v->scale)); /* max bytes per second */
  put32(fp, 0);                                         /* padding granularity */
  put32(fp, 0x10);                                      /* AVIF_HASINDEX */
  put32(fp, v->frames);
  put32(fp, 0);                                         /* initial frames */
  put32(fp, 1);                                         /* streams */
  put32(fp, v->frame_bytes);                            /* suggested buffer size */
  put32(fp, v->width); put32(fp, v->height);
  put32(fp, 0); put32(fp, 0); put32(fp, 0); put32(fp, 0);
  tag(fp, "LIST"); put32(fp, 4 + (8 + 56) + (8 + 40)); tag(fp, "strl");
  tag(fp, "strh"); put32(fp, 56);
  tag(fp, "vids"); tag(fp, "DIB ");
  put32(fp, 0);                                         /* flags */
  put16(fp, 0); put16(fp, 0);                           /* priority, language */
  put32(fp, 0);                                         /* initial frames */
  put32(fp, v->scale); put32(fp, v->rate);              /* rate / scale = frames per second */
  put32(fp, 0);                                         /* start */
  put32(fp, v->frames);                                 /* length */
  put32(fp, v->frame_bytes);                            /* suggested buffer size */
  put32(fp, 0xffffffffu);                               /* quality: default */
  put32(fp, 0);                                         /* sample size: varies per chunk */
  put16(fp, 0); put16(fp, 0); put16(fp, (uint16_t)v->width); put16(fp, (uint16_t)v->height); /* frame rectangle */
  tag(fp, "strf"); put32(fp, 40);
  put32(fp, 40); put32(fp, v->width); put32(fp, v->height); /* positive height: bottom-up rows */
  put16(fp, 1); put16(fp, 24);
  put32(fp, 0);                                         /* BI_RGB */
  put32(fp, v->frame_bytes);
  put32(fp, 0); put32(fp, 0); put32(fp, 0); put32(fp, 0);
  v->movi_list_pos = ftell(fp);
  tag(fp, "LIST"); put32(fp, movi_bytes); tag(fp, "movi");
}
}  // namespace

extern "C" {

prs_video *prs_video_open(const char *path, unsigned width, unsigned height, double fps) {
  if (!width || !height || width > 65535u || height > 65535u || !(fps > 0.0)) return nullptr;
  FILE *fp = fopen(path, "wb");
  if (!fp) return nullptr;
  prs_video *v = new prs_video;
  v->fp = fp;
  v->width = width;
  v->height = height;
  v->row_bytes = (width * 3u + 3u) & ~3u;
  v->frame_bytes = v->row_bytes * height;
  v->scale = 1000;
  v->rate = (uint32_t)(fps * 1000.0 + 0.5);
  v->usec_per_frame = (uint32_t)(1e6 / fps + 0.5);
  v->row.assign(v->row_bytes, 0);
  write_header(v);
  return v;
}

int prs_video_write(prs_video *v, const unsigned char *bgr) {
  if (!v || v->failed) return -1;
  /* RIFF sizes are 32 bit: header + chunks + index must stay below 4 GiB */
  const uint64_t after = 4096ull + (uint64_t)(v->frames + 1) * (8ull + v->frame_bytes + 16ull);
  if (after >= 0xfff00000ull) return -1;
  FILE *fp = v->fp;
  v->offsets.push_back((uint32_t)(ftell(fp) - (v->movi_list_pos + 8)));
  tag(fp, "00db");
  put32(fp, v->frame_bytes);
  for (uint32_t r = 0; r < v->height; r++) { /* bottom-up in the file */
    memcpy(v->row.data(), bgr + (size_t)(v->height - 1 - r) * v->width * 3, (size_t)v->width * 3);
    if (fwrite(v->row.data(), 1, v->row_bytes, fp) != v->row_bytes) { v->failed = true; return -1; }
  }
  v->frames++;
  return 0;
}

int prs_video_close(prs_video *v) {
  if (!v) return -1;
  FILE *fp = v->fp;
  tag(fp, "idx1");
  put32(fp, v->frames * 16);
  for (uint32_t i = 0; i < v->frames; i++) {
    tag(fp, "00db");
    put32(fp, 0x10); /* AVIIF_KEYFRAME */
    put32(fp, v->offsets[i]);
    put32(fp, v->frame_bytes);
  }
  write_header(v); /* final counts */
  const bool bad = v->failed || ferror(fp);
  const int frames = (int)v->frames;
  const bool close_bad = fclose(fp) != 0;
  delete v;
  return (bad || close_bad) ? -1 : frames;
}

}  // extern "C"
