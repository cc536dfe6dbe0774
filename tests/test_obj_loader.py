"""OBJ ingestion of the host scene builder (jetpbrt::Scene::CreateTriangleMesh, reference scene.cc:49-64 /
shape.cc:23-68): exercised through a tiny C++ driver compiled against the host sources (no GPU)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

DRIVER = r"""
#include <cstdio>
#include "host/scene.h"
using namespace jetpbrt;
int main(int argc, char** argv) {
    Scene s("obj");
    // reference call shape: CreateTriangleMesh(file, flip_normal, flipHandedness, offset, scale)  (main.cc:94)
    auto mesh = s.CreateTriangleMesh(argv[1], true, true, Vec3(10, 20, 30), 2.f);
    int m = s.CreateMatteMaterial(Vec3(.5f, .5f, .5f));
    s.CreatePrimitives(mesh, m);
    const jpbrt_scene_desc* d = s.Desc();
    printf("%d %d\n", (int)mesh.size(), d->n_primitives);
    for (int i = 0; i < d->n_shapes; ++i) {
        const jpbrt_shape& sh = d->shapes[i];
        printf("%d %d", sh.type, sh.flip_normal);
        for (int k = 0; k < 3; ++k) printf(" %g %g %g", sh.p[k][0], sh.p[k][1], sh.p[k][2]);
        printf("\n");
    }
    auto missing = s.CreateTriangleMesh("scene\\does\\not\\exist.obj");
    printf("missing %d\n", (int)missing.size());
    return 0;
}
"""

OBJ = """# quad + triangle, mixed index styles, negative index
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
vt 0 0
vn 0 0 1
f 1 2 3 4
f 1/1/1 2/1/1 5/1/1
f -5//1 -4//1 -1//1
"""


def test_obj_mesh_is_loaded_with_the_reference_transform(tmp_path):
    (tmp_path / "driver.cc").write_text(DRIVER)
    sub = tmp_path / "scene" / "m"
    sub.mkdir(parents=True)
    (sub / "mesh.obj").write_text(OBJ)
    exe = tmp_path / "driver"
    pk = ROOT / "jet-pbrt_b200"
    subprocess.run(["g++", "-std=gnu++17", "-O1", "-I", str(pk), str(tmp_path / "driver.cc"), str(pk / "host" / "scene.cc"),
                    "-o", str(exe)], check=True)
    # the reference spells paths with backslashes (main.cc:34); they must work here too
    arg = str(tmp_path) + "\\scene\\m\\mesh.obj"
    out = subprocess.run([str(exe), arg], capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "4 4"  # quad fan-triangulated (2) + 2 triangles
    rows = [list(map(float, l.split())) for l in out[1:5]]
    for r in rows:
        assert r[0] == 0 and r[1] == 1  # triangles, flip_normal kept
    # LoadTriangleMesh order (shape.cc:48-62): z negated, then * scale, then + offset
    assert rows[0][2:] == [10, 20, 30, 12, 20, 30, 12, 22, 30]
    assert rows[1][2:] == [10, 20, 30, 12, 22, 30, 10, 22, 30]
    assert rows[2][2:] == [10, 20, 30, 12, 20, 30, 10, 20, 28]   # (0,0,1) -> z = -1 -> -2 + 30
    assert rows[3][2:] == rows[2][2:]                            # negative indices name the same vertices
    assert out[5].startswith("load triangle mesh failed") and out[6] == "missing 0"  # failure prints and yields an empty mesh (shape.cc:28-32)
