#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- mechanical GLSL 4.50 -> C++ rewrite of the reference's shaders, so that g++ compiles the
reference's OWN shader text (read from /root/reference at build time, never copied into this repo) against glm, the
GLSL-semantics vector library the reference itself vendors.  oracle/Makefile target `refshaders` runs this into a
scratch directory under /tmp (removed after the compile) and builds oracle/_ref/libmeteoros_refshaders.so; oracle/glsl_rt.h supplies what a Vulkan device would: samplers, storage
images, gl_GlobalInvocationID, the dispatch loop.

The rewrite is token-level and changes no expression:
  * `#version`, `#extension`, `layout(local_size...) in;` are dropped;
  * `layout(...) uniform Block { fields } name;` becomes `struct Block_blk { fields } name;`; a block without an
    instance name becomes its fields as plain globals (GLSL scoping);
  * `layout(...) uniform [readonly|writeonly] image2D|sampler2D|sampler3D name;` becomes `image2D name;` etc.;
  * fragment-stage `layout(location = N) in|out T name;` becomes `thread_local T name;`;
  * parameter qualifiers: `in T x` -> `T x`, `out T x` / `inout T x` -> `T& x`;
  * floating literals get an `f` suffix (GLSL literals are 32-bit; C++ ones would be double);
  * r-value swizzles `.xyz` become glm's `.xyz()`; swizzles that are assigned to are left to glm's l-value proxies;
  * a declaration without an initialiser, `T a, b;`, becomes `T a{}, b{};`: GLSL leaves such a variable undefined (the
    shaders do read one -- raySphereIntersection returns isect.t unwritten when the ray misses); canonical value 0;
  * `void main()` becomes `void shader_main()`.
Everything else -- every arithmetic expression, branch, loop and constant -- is compiled exactly as the reference wrote it.

`--revive-weather` (cloudRayMarch.comp only) additionally un-comments the weather-map block the reference carries as
dead code (:517-524), takes the coverage from `weather_data.r` as the trailing comment of :529 says, and scales the
weather sample point by the global `mt_weather_scale` -- the semantics of MtTuning.use_weather."""
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")
SWIZZLE = re.compile(r"\.([xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})\b(?!\s*(?:[-+*/]?=(?!=))|\s*\()")
BLOCK = re.compile(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*(\w*)\s*;", re.S)
OPAQUE = re.compile(r"layout\s*\([^)]*\)\s*uniform\s+(?:(?:readonly|writeonly|coherent|restrict)\s+)*(image2D|sampler2D|sampler3D)\s+(\w+)\s*;")
TYPES = r"float|int|uint|bool|vec2|vec3|vec4|ivec2|ivec3|ivec4|uvec2|uvec3|mat3|mat4"
STAGE_IO = re.compile(r"layout\s*\(\s*location\s*=\s*\d+\s*\)\s*(?:in|out)\s+(\w+)\s+(\w+)\s*;")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def revive_weather(src: str) -> str:
    dead = r"^(\s*)//\s*(vec2 weatherSamplePoint|vec3 weather_data|float cloudType|float densityHeightGradient|baseCloud \*= densityHeightGradient)"
    src, n = re.subn(dead, r"\1\2", src, flags=re.M)
    assert n == 5, n
    src, n = re.subn(r"float cloud_coverage = 0\.6;// weather_data\.r;", "float cloud_coverage = weather_data.r;", src)
    assert n == 1, n
    src, n = re.subn(r"= unskewedSamplePoint\.xz;", "= unskewedSamplePoint.xz * mt_weather_scale;", src)
    assert n == 1, n
    return src


def rewrite(src: str, weather: bool = False) -> str:
    src = src.lstrip("\ufeff")
    if weather:
        src = revive_weather(src)
    src = strip_comments(src)
    src = re.sub(r"^\s*#\s*(version|extension)[^\n]*", "", src, flags=re.M)
    src = re.sub(r"layout\s*\(\s*local_size[^)]*\)\s*in\s*;", "", src)

    def block(m):
        name, fields, inst = m.group(1), m.group(2), m.group(3)
        return f"struct {name}_blk {{{fields}}} {inst};" if inst else fields

    src = BLOCK.sub(block, src)
    src = OPAQUE.sub(r"\1 \2;", src)
    src = STAGE_IO.sub(r"thread_local \1 \2;", src)
    # parameter qualifiers (only ever inside parameter lists in these shaders)
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"(?<=[(,])\s*in\s+(?=\w+\s+\w+)", " ", src)
    src = FLOAT_LIT.sub(r"\1f", src)
    src = SWIZZLE.sub(r".\1()", src)
    structs = re.findall(r"\bstruct\s+(\w+)", src)
    decl = re.compile(r"(?<![\w&,(])(" + "|".join([TYPES] + structs) + r")\s+(\w+(?:\s*,\s*\w+)*)\s*;")
    src = decl.sub(lambda m: m.group(1) + " " + ", ".join(n.strip() + "{}" for n in m.group(2).split(",")) + ";", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    return src


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    shader, out = args[0], args[1]
    with open(shader, encoding="utf-8-sig", errors="replace") as f:
        text = rewrite(f.read(), weather="--revive-weather" in sys.argv)
    with open(out, "w") as f:
        f.write(f"// GENERATED by oracle/glsl2cpp.py from {shader} -- scratch build input, never to be committed.\n")
        f.write(text)
