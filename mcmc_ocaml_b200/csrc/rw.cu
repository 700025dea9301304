// rw.cu -- Read_write (read_write.ml): the text format every bin/ tool of the
// reference uses to move samples between runs (bin/evidence_tool.ml:44,
// bin/harmonic_evidence.ml:39).  Host-only code (no kernels): it sits on
// either side of the GPU path, not on it.
//
// Format (read_write.ml:19-29): one sample per line, every coordinate as "%g "
// (C printf semantics, 6 significant digits -- lossy), then "%g %g\n" for
// log_likelihood and log_prior.  Nested output (read_write.ml:60-67): a first
// line "%g %g\n" (log_ev, log_dev), then the sample fields followed by the log
// weight.  `precision` = 0 writes the reference's "%g"; 17 writes "%.17g",
// which the reference's reader parses unchanged and which round-trips exactly.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcmc_gpu.h"

namespace {

struct File {
  FILE *f = nullptr; bool own = false;
  File(const char *path, const char *mode, FILE *dflt) {
    if (!path || strcmp(path, "-") == 0) f = dflt;
    else { f = fopen(path, mode); own = true; }
  }
  ~File() { if (own && f) fclose(f); }
};

void put(FILE *f, double x, int precision) {
  if (precision > 0) fprintf(f, "%.*g", precision, x); else fprintf(f, "%g", x);
}

// Scanf.bscanf " %g " repeatedly until End_of_file (read_write.ml:33-39): every
// whitespace-separated float of the line
bool parse_line(const char *s, std::vector<double> &out) {
  out.clear();
  for (;;) {
    while (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n') ++s;
    if (!*s) return true;
    char *end = nullptr;
    errno = 0;
    const double v = strtod(s, &end);
    if (end == s) return false;          // Scanf failure
    out.push_back(v);
    s = end;
  }
}

bool read_line(FILE *f, std::string &line) {
  line.clear();
  int c;
  bool any = false;
  while ((c = fgetc(f)) != EOF) { any = true; if (c == '\n') break; line.push_back((char)c); }
  return any;
}

int read_table(FILE *f, int tail, std::vector<double> &rows, int64_t *n, int32_t *D) {
  std::string line; std::vector<double> v;
  int64_t count = 0; int width = -1;
  while (read_line(f, line)) {
    if (!parse_line(line.c_str(), v)) return MG_EFAIL;                 // Scanf.Scan_failure
    if ((int)v.size() < tail) return MG_EINVAL;                       // Array.sub with a negative length
    if (width < 0) width = (int)v.size();
    else if ((int)v.size() != width) return MG_EINVAL;                // ragged input: not one float array type
    rows.insert(rows.end(), v.begin(), v.end());
    ++count;
  }
  *n = count; *D = width < 0 ? 0 : width - tail;
  return MG_OK;
}

}  // namespace

extern "C" void mg_free_host(void *p) { free(p); }

// Read_write.write (read_write.ml:19-29).  rows: [n][D+2].  path "-" = stdout.
extern "C" int mg_write_samples(const char *path, const double *rows, int64_t n, int32_t D, int32_t precision) {
  if (!rows && n > 0) return MG_EINVAL;
  File out(path, "w", stdout);
  if (!out.f) return MG_EFAIL;                                         // Sys_error
  for (int64_t i = 0; i < n; ++i) {
    const double *r = rows + i * (D + 2);
    for (int d = 0; d < D; ++d) { put(out.f, r[d], precision); fputc(' ', out.f); }
    put(out.f, r[D], precision); fputc(' ', out.f); put(out.f, r[D + 1], precision); fputc('\n', out.f);
  }
  return ferror(out.f) ? MG_EFAIL : MG_OK;
}

// Read_write.read (read_write.ml:31-58).  *rows is malloc'ed [n][D+2] (mg_free_host).  path "-" = stdin.
extern "C" int mg_read_samples(const char *path, double **rows, int64_t *n, int32_t *D) {
  if (!rows || !n || !D) return MG_EINVAL;
  *rows = nullptr; *n = 0; *D = 0;
  File in(path, "r", stdin);
  if (!in.f) return MG_EFAIL;
  std::vector<double> v;
  const int rc = read_table(in.f, 2, v, n, D);
  if (rc) return rc;
  if (!v.empty()) {
    *rows = (double *)malloc(v.size() * sizeof(double));
    if (!*rows) return MG_ENOMEM;
    memcpy(*rows, v.data(), v.size() * sizeof(double));
  }
  return MG_OK;
}

// Read_write.write_nested (read_write.ml:60-67)
extern "C" int mg_write_nested(const char *path, double log_ev, double log_dev, const double *rows, const double *logw,
                               int64_t n, int32_t D, int32_t precision) {
  if ((!rows || !logw) && n > 0) return MG_EINVAL;
  File out(path, "w", stdout);
  if (!out.f) return MG_EFAIL;
  put(out.f, log_ev, precision); fputc(' ', out.f); put(out.f, log_dev, precision); fputc('\n', out.f);
  for (int64_t i = 0; i < n; ++i) {
    const double *r = rows + i * (D + 2);
    for (int d = 0; d < D + 2; ++d) { put(out.f, r[d], precision); fputc(' ', out.f); }
    put(out.f, logw[i], precision); fputc('\n', out.f);
  }
  return ferror(out.f) ? MG_EFAIL : MG_OK;
}

// Read_write.read_nested (read_write.ml:69-101).  *rows [n][D+2], *logw [n], malloc'ed.
extern "C" int mg_read_nested(const char *path, double *log_ev, double *log_dev, double **rows, double **logw,
                              int64_t *n, int32_t *D) {
  if (!log_ev || !log_dev || !rows || !logw || !n || !D) return MG_EINVAL;
  *rows = *logw = nullptr; *n = 0; *D = 0;
  File in(path, "r", stdin);
  if (!in.f) return MG_EFAIL;
  std::string line; std::vector<double> v;
  if (!read_line(in.f, line)) return MG_EFAIL;                         // End_of_file on the header (input_line)
  if (!parse_line(line.c_str(), v) || v.size() < 2) return MG_EFAIL;
  *log_ev = v[0]; *log_dev = v[1];
  std::vector<double> t;
  int32_t Dw = 0;
  const int rc = read_table(in.f, 3, t, n, &Dw);
  if (rc) return rc;
  *D = Dw;
  if (*n > 0) {
    const int W = Dw + 3;
    *rows = (double *)malloc((size_t)*n * (Dw + 2) * sizeof(double));
    *logw = (double *)malloc((size_t)*n * sizeof(double));
    if (!*rows || !*logw) return MG_ENOMEM;
    for (int64_t i = 0; i < *n; ++i) {
      memcpy(*rows + i * (Dw + 2), t.data() + i * W, (Dw + 2) * sizeof(double));
      (*logw)[i] = t[i * W + Dw + 2];
    }
  }
  return MG_OK;
}
