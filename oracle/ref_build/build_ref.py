#!/usr/bin/env python3
"""Build oracle/_ref/ from the reference's own sources where they lie under /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Nothing is copied into the repository: every derived
file (pre-processed shader text, shim headers with Windows-style names, objects, .so) goes to
oracle/_ref/, which is git-ignored but travels to the GPU box with gpurun.

  libref_shaders.so   the reference's UNMODIFIED fragment shaders compiled as C++ through glsl_shim.h
                      (float literals get an `f` suffix, `#extension` lines are dropped, file-scope
                      variables become thread_local so rows can run in parallel; nothing else changes)
  libref_host.so      the reference's Mesh / SceneLoader / OBJLoader / ShadowVolume /
                      UniformSampledLightSource translation units + its vendored GLM 0.9.3.1, compiled
                      against stub GL/OpenCV headers (ref_host_glue.cpp exports a C API over them)

Usage: python oracle/ref_build/build_ref.py [--reference /root/reference]
"""
import argparse
import itertools
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.normpath(os.path.join(HERE, "..", "_ref"))
GEN = os.path.join(OUT, "gen")

SHADERS = {
    # name -> path under the reference
    "shadow": "ShadowMapping/Shaders/Shadow.frag",
    "nonconservative": "ShadowMapping/Shaders/RBSM/NonConservativeSMSR.frag",
    "conservative": "ShadowMapping/Shaders/RBSM/ConservativeSMSR.frag",
    "filtered": "ShadowMapping/Shaders/RBSM/FilteredRBSM.frag",
    "plausible": "SoftShadowMapping/Shaders/SoftShadow/PlausibleSoftShadow.frag",
    "accurate": "SoftShadowMapping/Shaders/SoftShadow/AccurateSoftShadow.frag",
    "phong": "ShadowMapping/Shaders/GBuffer/PhongShading.frag",
    "gbuffer": "ShadowMapping/Shaders/GBuffer/GBuffer.frag",      # computeFragmentColor: texture select / vertex colour (third MRT)
    "rbssm": "SoftShadowMapping/Shaders/SoftShadow/RBSSM.frag",
    "meanfilter": "ShadowMapping/Shaders/Filter/MeanFilter.frag",
    # moment shadow maps: the light-view fragment programs and the separable blurs
    "moments": "ShadowMapping/Shaders/ShadowMap/Moments.frag",
    "exponential": "ShadowMapping/Shaders/ShadowMap/Exponential.frag",
    "expmoments": "ShadowMapping/Shaders/ShadowMap/ExponentialMoments.frag",
    "gaussian": "ShadowMapping/Shaders/Filter/GaussianFilter.frag",
    "loggaussian": "ShadowMapping/Shaders/Filter/LogGaussianFilter.frag",
}


def cxx():
    return "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"


def gen_swizzles():
    sets = ["xyzw", "rgba", "stpq"]
    for n in (2, 3, 4):
        lines = []
        for letters in sets:
            ls = letters[:n]
            for k in (2, 3, 4):
                for combo in itertools.product(range(n), repeat=k):
                    name = "".join(ls[i] for i in combo)
                    args = ", ".join(str(i) for i in combo)
                    lines.append(f"    Sw{k}<{n}, {args}> {name};")
        with open(os.path.join(GEN, f"swizzles_vec{n}.inc"), "w") as f:
            f.write("\n".join(lines) + "\n")


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
VARYING = re.compile(r"^\s*varying\s+vec[34]\s+(\w+)\s*;")
VARYING3 = re.compile(r"^\s*varying\s+vec3\s+(\w+)\s*;")
# which gl_FragData[] element the runner returns (default 0)
OUT_INDEX = {"gbuffer": 2}
UNIFORM = re.compile(r"^\s*uniform\s+(\w+)\s+(\w+)\s*(?:\[\s*(\d+)\s*\])?\s*;")
PLAIN_GLOBAL = re.compile(r"^(float|int|bool|vec[234]|mat[34])\s+\w+\s*;")


# Type leniencies of the GLSL compiler the reference was written against (NVIDIA's accepts a float where a vec4 return
# type is declared and assigns the vec4 back to a float through .x): stated here as the equivalent declaration.
DECL_FIXES = {
    "vec4 revectorizationBasedShadowMappingSmoothing(vec4 normalizedShadowCoord)":
        "float revectorizationBasedShadowMappingSmoothing(vec4 normalizedShadowCoord)",     # RBSSM.frag:1341 (every return is a float)
    # LogGaussianFilter.frag:10 declares float for an expression of four identical components that the caller wraps in vec4()
    "float log_conv ( float x0, vec4 X, float y0, vec4 Y )": "vec4 log_conv ( float x0, vec4 X, float y0, vec4 Y )",
}


def transform(src):
    """Literal suffixes, drop #extension, thread_local file-scope variables. Returns (text, uniforms)."""
    out, uniforms, depth = [], [], 0
    varyings = []
    vec3s = set()
    for a, b in DECL_FIXES.items():
        src = src.replace(a, b)
    for line in src.splitlines():
        if line.lstrip().startswith("#extension"):
            out.append("// " + line)
            continue
        code = line.split("//")[0]
        m = UNIFORM.match(code)
        if m:
            uniforms.append((m.group(1), m.group(2), int(m.group(3)) if m.group(3) else 0))
        mv = VARYING.match(code)
        if mv:
            varyings.append(mv.group(1))
            if VARYING3.match(code):
                vec3s.add(mv.group(1))
        new = FLOAT_LIT.sub(lambda mm: mm.group(1) + "f", line)
        if depth == 0 and PLAIN_GLOBAL.match(code):
            new = "thread_local " + new
        depth += code.count("{") - code.count("}")
        out.append(new)
    return "\n".join(out) + "\n", uniforms, [(v, v in vec3s) for v in varyings]


RUNNER = r"""
struct RefBinding { const char* name; const void* data; uint64_t size; };

// per-pixel inputs of programs that are not full-screen passes: `varying vec4` images ("varying:<name>", float4[H][W]) and
// the values dFdx / dFdy return for this fragment ("dFdx", "dFdy", float[H][W]; the harness supplies the derivative, the
// shader's own arithmetic runs unmodified)
static thread_local const float* vary_img[8];
static thread_local const float* ddx_img;
static thread_local const float* ddy_img;
static int bind_all(const RefBinding* b, int n) {
  int bound = 0;
  for (int k = 0; k < 8; k++) vary_img[k] = nullptr;
  ddx_img = nullptr; ddy_img = nullptr;
  for (int i = 0; i < n; i++) {
    if (!strcmp(b[i].name, "dFdx")) { ddx_img = (const float*)b[i].data; continue; }
    if (!strcmp(b[i].name, "dFdy")) { ddy_img = (const float*)b[i].data; continue; }
    %(VBINDS)s
    %(BINDS)s
  }
  return bound;
}
static inline void set_varyings(size_t o) {
  if (ddx_img) glsl::g_dfdx = ddx_img[o];
  if (ddy_img) glsl::g_dfdy = ddy_img[o];
  %(VSETS)s
}
}  // namespace shader_%(NAME)s

extern "C" int ref_%(NAME)s_run(const shader_%(NAME)s::RefBinding* b, int nb, int W, int H, int x0, int y0, int x1,
                                int y1, float* out0) {
  using namespace shader_%(NAME)s;
#pragma omp parallel
  {
    bind_all(b, nb);
#pragma omp for schedule(dynamic, 4)
    for (int j = y0; j < y1; j++)
      for (int i = x0; i < x1; i++) {
        // Shadow.vert:7-8: f_texcoord = texcoord*0.5+0.5 with texcoord the quad's NDC position
        float nx = ((float)i + 0.5f) / (float)W * 2.0f - 1.0f, ny = ((float)j + 0.5f) / (float)H * 2.0f - 1.0f;
        %(TEXCOORD)sf_texcoord = glsl::vec2(nx * 0.5f + 0.5f, ny * 0.5f + 0.5f);
        gl_discarded = false;
        set_varyings((size_t)j * W + i);
        float* o = out0 + 4 * ((size_t)j * W + i);
        gl_FragData[0] = glsl::vec4(o[0], o[1], o[2], o[3]);
        shader_main();
        if (gl_discarded) continue;
        o[0] = gl_FragData[%(OUTIDX)s].x; o[1] = gl_FragData[%(OUTIDX)s].y; o[2] = gl_FragData[%(OUTIDX)s].z; o[3] = gl_FragData[%(OUTIDX)s].w;
      }
  }
  return 0;
}
"""


def gen_shader_tu(name, ref_root):
    path = os.path.join(ref_root, SHADERS[name])
    with open(path, "r", errors="replace") as f:
        text, uniforms, varyings = transform(f.read())
    binds = []
    for ty, nm, arr in uniforms:
        binds.append(
            f'if (!strcmp(b[i].name, "{nm}")) {{ memcpy((void*)&{nm}, b[i].data, b[i].size < sizeof({nm}) ? b[i].size : sizeof({nm})); bound++; continue; }}'
        )
    tu = [
        '#include "glsl_shim.h"',
        "#include <omp.h>",
        f"namespace shader_{name} {{",
        "using namespace glsl;",
        "thread_local vec4 gl_FragData[4];",
        "thread_local bool gl_discarded;",
        "thread_local bool gl_FrontFacing;",
        "thread_local LightModelProducts gl_FrontLightModelProduct;",
        "#define gl_FragColor gl_FragData[0]",
        "#define uniform thread_local",
        "#define varying thread_local",
        "#define main shader_main",
        "#define discard do { gl_discarded = true; return; } while (0)",
        f'#line 1 "{path}"',
        text,
        "#undef uniform",
        "#undef varying",
        "#undef main",
        "#undef discard",
        RUNNER % {"NAME": name, "BINDS": "\n    ".join(binds), "OUTIDX": str(OUT_INDEX.get(name, 0)),
                  "TEXCOORD": "" if "f_texcoord" in text else "(void)nx; (void)ny; // no f_texcoord in this program: ",
                  "VBINDS": "\n    ".join(
                      f'if (!strcmp(b[i].name, "varying:{v}")) {{ vary_img[{k}] = (const float*)b[i].data; continue; }}'
                      for k, (v, _) in enumerate(varyings)),
                  "VSETS": "\n  ".join(
                      (f"if (vary_img[{k}]) {v} = glsl::vec3(vary_img[{k}][4 * o], vary_img[{k}][4 * o + 1], vary_img[{k}][4 * o + 2]);" if is3 else
                       f"if (vary_img[{k}]) {v} = glsl::vec4(vary_img[{k}][4 * o], vary_img[{k}][4 * o + 1], vary_img[{k}][4 * o + 2], vary_img[{k}][4 * o + 3]);")
                      for k, (v, is3) in enumerate(varyings))},
    ]
    out = os.path.join(GEN, f"shader_{name}.cpp")
    with open(out, "w") as f:
        f.write("\n".join(tu))
    return out


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_shaders(ref_root):
    gen_swizzles()
    tus = [gen_shader_tu(n, ref_root) for n in SHADERS]
    flags = ["-O2", "-std=gnu++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-w",
             "-I", HERE, "-I", GEN]
    objs = []
    procs = []
    for tu in tus:
        obj = tu[:-4] + ".o"
        objs.append(obj)
        cmd = [cxx()] + flags + ["-c", tu, "-o", obj]
        print("+", " ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("shader compile failed")
    run([cxx(), "-shared", "-fopenmp", "-o", os.path.join(OUT, "libref_shaders.so")] + objs)


# ------------------------------------------------------------------------------------------------ host
HOST_STUBS = {
    # file name (may contain a backslash: the reference uses Windows include paths) -> content
    "GL/glut.h": r"""
#pragma once
// stub: only the typedefs/constants/entry points OBJLoader.cpp names; nothing is ever drawn
typedef unsigned int GLuint; typedef int GLint; typedef float GLfloat; typedef void GLvoid;
typedef unsigned int GLenum; typedef unsigned char GLubyte; typedef double GLdouble; typedef int GLsizei;
typedef unsigned char GLboolean;
#define GL_TRIANGLES 4
#define GL_TRUE 1
#define GL_FALSE 0
#define GL_FRONT_AND_BACK 0x0408
#define GL_AMBIENT 0x1200
#define GL_DIFFUSE 0x1201
#define GL_SPECULAR 0x1202
#define GL_SHININESS 0x1601
#define GL_COLOR_MATERIAL 0x0B57
#define GL_COMPILE 0x1300
static inline void glEnable(GLenum) {} static inline void glDisable(GLenum) {}
static inline void glMaterialfv(GLenum, GLenum, const GLfloat*) {} static inline void glMaterialf(GLenum, GLenum, GLfloat) {}
static inline void glColor3fv(const GLfloat*) {} static inline void glBegin(GLenum) {} static inline void glEnd() {}
static inline void glNormal3fv(const GLfloat*) {} static inline void glTexCoord2fv(const GLfloat*) {}
static inline void glVertex3fv(const GLfloat*) {} static inline GLuint glGenLists(GLsizei) { return 0; }
static inline void glNewList(GLuint, GLenum) {} static inline void glEndList() {}
""",
    "opencv2\\opencv.hpp": r"""
#pragma once
// stub: textures only affect colour (SURVEY C4); Image.cpp is replaced by ref_host_glue.cpp's no-op Image
#include <string>
namespace cv { struct Mat { unsigned char* data = nullptr; int rows = 0, cols = 0; bool empty() const { return true; } int channels() const { return 3; } };
  inline Mat imread(const std::string&, int = 1) { return Mat(); } }
""",
    "IO\\SceneLoader.h": '#include "IO/SceneLoader.h"\n',
    "Scene\\LightSource\\UniformSampledLightSource.h": '#include "Scene/LightSource/UniformSampledLightSource.h"\n',
    "Scene\\LightSource\\LightSource.h": '#include "Scene/LightSource/LightSource.h"\n',
}


def build_host(ref_root):
    stub = os.path.join(GEN, "stubs")
    for name, content in HOST_STUBS.items():
        p = os.path.join(stub, name) if "\\" not in name else os.path.join(stub, name)
        d = os.path.dirname(p) if "\\" not in name else stub
        os.makedirs(d, exist_ok=True)
        with open(p, "w") as f:
            f.write(content)
    sm = os.path.join(ref_root, "ShadowMapping")
    ssm = os.path.join(ref_root, "SoftShadowMapping")
    sv = os.path.join(ref_root, "ShadowVolumes")
    flags = ["-O1", "-std=gnu++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fpermissive", "-w",
             # MSVC's headers pull these in transitively; gcc's do not
             "-include", "cstring", "-include", "cstdlib", "-include", "cstdio", "-include", "cmath",
             "-I", stub]
    objs = []

    def cc(src, inc, tag, extra=()):
        obj = os.path.join(GEN, f"host_{tag}.o")
        run([cxx()] + flags + ["-I", inc] + list(extra) + ["-c", src, "-o", obj])
        objs.append(obj)

    # ShadowMapping flavour of Mesh/SceneLoader/OBJLoader (the SV/SSM copies differ only in `h*`/`d` keys)
    cc(os.path.join(sm, "src/Mesh.cpp"), os.path.join(sm, "include"), "mesh")
    cc(os.path.join(sm, "src/IO/SceneLoader.cpp"), os.path.join(sm, "include"), "sceneloader")
    cc(os.path.join(sm, "src/IO/OBJLoader.cpp"), os.path.join(sm, "include"), "objloader")
    # ShadowVolume.cpp uses the SV Mesh.h (same class layout for the members it touches; compiled against SM's
    # header so that one Mesh type is shared with the loader above)
    cc(os.path.join(sv, "src/ShadowVolume.cpp"), os.path.join(sm, "include"), "shadowvolume",
       ["-I", os.path.join(sv, "include")])
    cc(os.path.join(ssm, "src/Scene/LightSource/UniformSampledLightSource.cpp"), os.path.join(ssm, "include"), "uniformlight")
    cc(os.path.join(ssm, "src/Scene/LightSource/LightSource.cpp"), os.path.join(ssm, "include"), "lightsource")
    cc(os.path.join(HERE, "ref_host_glue.cpp"), os.path.join(sm, "include"), "glue",
       ["-I", os.path.join(sv, "include"), "-I", os.path.join(ssm, "include"), "-DSSM_INCLUDE"])
    run([cxx(), "-shared", "-o", os.path.join(OUT, "libref_host.so")] + objs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", choices=["shaders", "host"], default=None)
    a = ap.parse_args()
    if not os.path.isdir(a.reference):
        print(f"reference tree {a.reference} not present: keeping any prebuilt oracle/_ref as is")
        return 0
    os.makedirs(GEN, exist_ok=True)
    if a.only in (None, "shaders"):
        build_shaders(a.reference)
    if a.only in (None, "host"):
        build_host(a.reference)
    return 0


if __name__ == "__main__":
    sys.exit(main())
