// obj_loader.cpp — see obj_loader.h.  Two passes over an in-memory copy of the file with a cursor that
// behaves like the C stdio calls the reference's reader makes (fscanf "%s"/"%d"/"%f", fgets(buf,128)).
#include "obj_loader.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sgh {
namespace {

struct Cursor {
  const char* p; const char* end;
  bool eof() const { return p >= end; }
  void skip_ws() { while (p < end && isspace((unsigned char)*p)) p++; }
  // fscanf(file, "%s", buf): false at EOF
  bool token(std::string* out) {
    skip_ws();
    if (p >= end) return false;
    const char* s = p;
    while (p < end && !isspace((unsigned char)*p)) p++;
    out->assign(s, p - s);
    return true;
  }
  // fgets(buf, 128, file): at most 127 characters, stops after '\n'; returns the text read
  std::string eat_line() {
    const char* s = p;
    int n = 0;
    while (p < end && n < 127) { char c = *p++; n++; if (c == '\n') break; }
    return std::string(s, p - s);
  }
  // "%d": skips white space, optional sign, digits. false (nothing consumed but ws / sign) on mismatch
  bool integer(int* v) {
    skip_ws();
    const char* s = p;
    if (p < end && (*p == '+' || *p == '-')) p++;
    if (p >= end || !isdigit((unsigned char)*p)) { p = (p > s && (p >= end || !isdigit((unsigned char)*p))) ? p : s; return false; }
    long long acc = 0;
    bool neg = (*s == '-');
    while (p < end && isdigit((unsigned char)*p)) { acc = acc * 10 + (*p - '0'); if (acc > 0x7fffffffLL) acc = 0x7fffffffLL; p++; }
    *v = (int)(neg ? -acc : acc);
    return true;
  }
  bool literal(char c) { if (p < end && *p == c) { p++; return true; } return false; }
  // "%f"
  bool real(float* v) {
    skip_ws();
    if (p >= end) return false;
    // strtof needs a NUL-terminated string: copy the candidate token
    char tmp[128];
    size_t n = 0;
    const char* q = p;
    while (q < end && n < sizeof(tmp) - 1 && !isspace((unsigned char)*q)) tmp[n++] = *q++;
    tmp[n] = 0;
    char* e = nullptr;
    float f = strtof(tmp, &e);
    if (e == tmp) return false;
    p += (e - tmp);
    *v = f;
    return true;
  }
};

enum FaceKind { F_V, F_VT, F_VN, F_VTN };

// the format probe the reference applies to the first vertex token of a face
FaceKind classify(const std::string& tok, int* v, int* t, int* n) {
  if (tok.find("//") != std::string::npos) { sscanf(tok.c_str(), "%d//%d", v, n); return F_VN; }
  if (sscanf(tok.c_str(), "%d/%d/%d", v, t, n) == 3) return F_VTN;
  if (sscanf(tok.c_str(), "%d/%d", v, t) == 2) return F_VT;
  sscanf(tok.c_str(), "%d", v);
  return F_V;
}

// fscanf(file, "%d//%d" | "%d/%d/%d" | "%d/%d" | "%d"): number of fields assigned
int scan_vertex(Cursor& c, FaceKind k, int* v, int* t, int* n) {
  int got = 0;
  if (!c.integer(v)) return 0;
  got = 1;
  switch (k) {
    case F_V: return got;
    case F_VN:
      if (!c.literal('/') || !c.literal('/')) return got;
      if (!c.integer(n)) return got;
      return 2;
    case F_VT:
      if (!c.literal('/')) return got;
      if (!c.integer(t)) return got;
      return 2;
    case F_VTN:
      if (!c.literal('/')) return got;
      if (!c.integer(t)) return got;
      got = 2;
      if (!c.literal('/')) return got;
      if (!c.integer(n)) return got;
      return 3;
  }
  return got;
}

int first_pass(Cursor c, ObjModel* m, std::string* err) {
  std::string buf;
  uint32_t nv = 0, nn = 0, nt = 0, ntri = 0;
  while (c.token(&buf)) {
    switch (buf[0]) {
      case '#': c.eat_line(); break;
      case 'v':
        if (buf.size() == 1) { c.eat_line(); nv++; }
        else if (buf[1] == 'n') { c.eat_line(); nn++; }
        else if (buf[1] == 't') { c.eat_line(); nt++; }
        else { if (err) *err = "readOBJ: unknown token \"" + buf + "\""; return -2; }
        break;
      case 'f': {
        int v = 0, n = 0, t = 0;
        std::string tok;
        if (!c.token(&tok)) break;
        FaceKind k = classify(tok, &v, &t, &n);
        scan_vertex(c, k, &v, &t, &n);
        scan_vertex(c, k, &v, &t, &n);
        ntri++;
        while (scan_vertex(c, k, &v, &t, &n) > 0) ntri++;
        break;
      }
      default: c.eat_line(); break;      // m(tllib), u(semtl), g(roup), anything else
    }
  }
  m->numvertices = nv; m->numnormals = nn; m->numtexcoords = nt; m->numtriangles = ntri;
  return 0;
}

inline uint32_t rel(int i, uint32_t count_plus_one) { return (uint32_t)(i < 0 ? i + (int)count_plus_one : i); }

void second_pass(Cursor c, ObjModel* m) {
  std::string buf;
  uint32_t nv = 1, nn = 1, nt = 1, ntri = 0;
  float dummy;
  auto T = [&](uint32_t i) -> ObjTriangle& { return m->triangles[i]; };
  while (c.token(&buf)) {
    switch (buf[0]) {
      case '#': c.eat_line(); break;
      case 'v':
        if (buf.size() == 1) {
          for (int k = 0; k < 3; k++) if (!c.real(nv <= m->numvertices ? &m->vertices[3 * nv + k] : &dummy)) break;
          nv++;
        } else if (buf[1] == 'n') {
          for (int k = 0; k < 3; k++) if (!c.real(nn <= m->numnormals ? &m->normals[3 * nn + k] : &dummy)) break;
          nn++;
        } else if (buf[1] == 't') {
          for (int k = 0; k < 2; k++) if (!c.real(nt <= m->numtexcoords ? &m->texcoords[2 * nt + k] : &dummy)) break;
          nt++;
        }
        break;
      case 'f': {
        int v = 0, n = 0, t = 0;
        std::string tok;
        if (!c.token(&tok)) break;
        FaceKind k = classify(tok, &v, &t, &n);
        if (ntri >= m->numtriangles) { c.eat_line(); break; }
        auto put = [&](uint32_t tri, int slot) {
          T(tri).vindices[slot] = rel(v, nv);
          if (k == F_VT || k == F_VTN) T(tri).tindices[slot] = rel(t, nt);
          if (k == F_VN || k == F_VTN) T(tri).nindices[slot] = rel(n, nn);
        };
        put(ntri, 0);
        scan_vertex(c, k, &v, &t, &n); put(ntri, 1);
        scan_vertex(c, k, &v, &t, &n); put(ntri, 2);
        ntri++;
        while (ntri < m->numtriangles + 1 && scan_vertex(c, k, &v, &t, &n) > 0) {
          if (ntri >= m->numtriangles) break;
          T(ntri).vindices[0] = T(ntri - 1).vindices[0]; T(ntri).vindices[1] = T(ntri - 1).vindices[2];
          T(ntri).tindices[0] = T(ntri - 1).tindices[0]; T(ntri).tindices[1] = T(ntri - 1).tindices[2];
          T(ntri).nindices[0] = T(ntri - 1).nindices[0]; T(ntri).nindices[1] = T(ntri - 1).nindices[2];
          put(ntri, 2);
          ntri++;
        }
        break;
      }
      default: c.eat_line(); break;
    }
  }
}

}  // namespace

int readOBJ(const std::string& filename, ObjModel* model, std::string* err) {
  FILE* f = fopen(filename.c_str(), "rb");
  if (!f) { if (err) *err = "readOBJ: can't open data file \"" + filename + "\""; return -1; }
  std::string data;
  char chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) data.append(chunk, n);
  fclose(f);
  // text-mode "\r\n" needs no special care: '\r' is white space to every scanner used here
  *model = ObjModel();
  Cursor c{data.data(), data.data() + data.size()};
  int rc = first_pass(c, model, err);
  if (rc) return rc;
  model->vertices.assign(3 * ((size_t)model->numvertices + 1), 0.0f);
  if (model->numnormals) model->normals.assign(3 * ((size_t)model->numnormals + 1), 0.0f);
  if (model->numtexcoords) model->texcoords.assign(2 * ((size_t)model->numtexcoords + 1), 0.0f);
  model->triangles.assign(model->numtriangles, ObjTriangle{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}});
  second_pass(c, model);
  return 0;
}

}  // namespace sgh
