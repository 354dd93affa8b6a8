"""oracle/glsl2cpp.py -- the token-level GLSL -> C++ rewrite behind oracle/_ref (the reference's shaders compiled from their
own text).  These tests pin each rewrite rule on small snippets, so that "no expression is touched" stays checkable on a
machine without the reference tree; tests/test_reference_shaders.py exercises the real shaders where it exists."""
import importlib.util
from pathlib import Path

import pytest

_spec = importlib.util.spec_from_file_location("glsl2cpp", Path(__file__).resolve().parents[1] / "oracle" / "glsl2cpp.py")
glsl2cpp = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(glsl2cpp)


def rw(src):
    return " ".join(glsl2cpp.rewrite(src).split())


@pytest.mark.parametrize("src,want", [
    ("float a = 1.0;", "float a = 1.0f;"),
    ("float a = 1. / 255.;", "float a = 1.f / 255.f;"),
    ("float a = .5 + 2.5E-6 - 3e5;", "float a = .5f + 2.5E-6f - 3e5f;"),
    ("float a = 0.06f + 1.0f;", "float a = 0.06f + 1.0f;"),                 # already suffixed: untouched
    ("int i = 16; x = haltonSeq1.x + v2.y;", "int i = 16; x = haltonSeq1.x + v2.y;"),  # integers and members: untouched
    ("#define SUN 0.999956676946448443553574619906976478926848692873900859324", "#define SUN 0.999956676946448443553574619906976478926848692873900859324f"),
])
def test_float_literals_become_single_precision(src, want):
    assert rw(src) == want


def test_swizzles():
    assert rw("vec3 e = -camera.eye.xyz;") == "vec3 e = -camera.eye.xyz();"
    assert rw("enc -= enc.yzww * vec2(1./255., 0.).xxxy;") == "enc -= enc.yzww() * vec2(1.f/255.f, 0.f).xxxy();"
    assert rw("point.xy += curl.xy * 0.5;") == "point.xy += curl.xy() * 0.5f;"       # assigned-to swizzle: glm's l-value proxy
    assert rw("if (a.xy == b.xy) c = d.rgb;") == "if (a.xy() == b.xy()) c = d.rgb();"  # == is a comparison, not an assignment
    assert rw("float x = v.x; float t = isect.t;") == "float x = v.x; float t = isect.t;"


def test_uniform_blocks_and_opaque_types():
    named = rw("layout (set = 2, binding = 0) uniform CameraUBO { mat4 view; vec4 eye; } camera;")
    assert named == "struct CameraUBO_blk { mat4 view{}; vec4 eye{}; } camera;"
    anon = rw("layout (set = 3, binding = 0) uniform TimeUBO { vec2 time; int frameCountMod16; };")
    assert anon == "vec2 time{}; int frameCountMod16{};"                             # members become globals (GLSL scoping)
    assert rw("layout (set = 0, binding = 0, rgba16f) uniform writeonly image2D img;") == "image2D img;"
    assert rw("layout (set = 1, binding = 0) uniform sampler3D s;") == "sampler3D s;"
    assert rw("layout(location = 0) in vec2 in_uv;") == "thread_local vec2 in_uv{};"
    assert rw("#version 450\n#extension GL_ARB_x : enable\nlayout (local_size_x = 32, local_size_y = 32) in;\nint a = 1;") == "int a = 1;"


def test_parameter_qualifiers_and_main():
    assert rw("float f(in int index, ivec2 dim) { return 0.0; }") == "float f( int index, ivec2 dim) { return 0.0f; }"
    assert rw("void g(in vec2 uv, inout vec4 v, out float d) {}") == "void g( vec2 uv, vec4& v, float& d) {}"
    assert rw("void main() { }") == "void shader_main() { }"


def test_uninitialised_declarations_read_zero():
    assert rw("struct Hit { vec3 p; bool valid; float t; }; Hit h; vec4 cmin, cmax, cavg; float x = 1.0; vec3 q = vec3(0.0);") == \
        "struct Hit { vec3 p{}; bool valid{}; float t{}; }; Hit h{}; vec4 cmin{}, cmax{}, cavg{}; float x = 1.0f; vec3 q = vec3(0.0f);"
    assert rw("vec3 f(vec3 a, float b);") == "vec3 f(vec3 a, float b);"               # parameters are not declarations


def test_comments_are_dropped_before_anything_else():
    assert rw("float a = 1.0; // was 2.0 * b.xy\n/* float c = 3.0; */ float d = 4.0;") == "float a = 1.0f; float d = 4.0f;"


def test_nothing_else_changes():
    body = "for (float t = s; t < e; t += h) { acc += mix(a, b, clamp(t * k, lo, hi)) * pow(d, vec3(g)); if (acc >= one) { break; } }"
    assert rw(body) == body


@pytest.mark.skipif(not Path("/root/reference/src/CloudScapes/shaders/cloudRayMarch.comp").exists(), reason="reference tree not present")
def test_revive_weather_touches_exactly_the_dead_block():
    text = Path("/root/reference/src/CloudScapes/shaders/cloudRayMarch.comp").read_text(encoding="utf-8-sig", errors="replace")
    plain, revived = glsl2cpp.rewrite(text).splitlines(), glsl2cpp.rewrite(text, weather=True).splitlines()
    assert len(plain) == len(revived)
    changed = [(a, b) for a, b in zip(plain, revived) if a != b]
    assert len(changed) == 6                                                          # five revived statements + the coverage line
    assert all(not a.strip() for a, _ in changed[:5])                                 # they were comments (blank after stripping)
    assert "weather_data.r" in changed[5][1] and "0.6" in changed[5][0]
