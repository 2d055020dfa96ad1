"""Scene ingestion, host side: the reference's XML scene description and its OBJ/MTL subset -> `SceneDesc`.

Mirrors:
  * `Parse.SceneState.GetSceneState` / `InitSceneState`   EngineCore/Scene/Scene.fs:28-271
  * `LoadObjModel`, `Face.ToHitable`, `ObjState`           EngineCore/Models/ObjModelLoader.fs:13-92,283-340
  * `LoadObjMtl`, `ProcessMaterial`                        EngineCore/Models/Obj_Mtl.fs:163-217

Same results on the files the reference can load, with its quirks kept where they decide the picture:
  * every `newmtl` becomes `Lambertian(Ka)` appended to the GLOBAL material table before the XML's own materials, and
    the XML's `material` ints index that global table (quirk Q8: Scene.fs:252,258-259, Obj_Mtl.fs:195-197);
  * a Shape re-creates the group's faces with the XML material, whatever `usemtl` said (Scene.fs:154-161);
  * 3 vertex references -> Triangle, 4 -> Rect(p0,p1,p2,p3); negative references count from the end
    (ObjModelLoader.fs:63-92); faces before any `g` go to group "default"; `usemtl` defaults to "white" -> index 0;
  * the light is the FIRST face of its group and must be a quad; its normal is trig1's (Scene.fs:194-197);
  * `InitSceneState` is handed XML *text*, not a file name (quirk Q12, Scene.fs:266);
  * the integrator is `PathIntegrator(bvh, 3, light)` (Scene.fs:304).
Fixed rather than reproduced (the reference's FParsec grammar stops silently at the first statement it cannot parse,
ObjModelLoader.fs:283-294: `vt`/`vn` lines hit the `v` rule, `o` has no rule): `vt`, `vn`, `o`, `s`, `usemap`,
`maplib` and unknown keywords are read and ignored, so spot / Renault load.  Faces with more than 4 references are
rejected like the reference's `assert(false)`.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from .scene import (AreaLight, PinholeCamera, SceneDesc, make_materials, make_prims, TRIANGLE, RECT, PATH_INTEGRATOR)


class IngestError(ValueError):
    pass


class ObjState:
    """ObjModelLoader.fs:17-53: vertices and the faces of every group, in file order."""

    def __init__(self):
        self.vertices = []
        self.groups = {"default": []}          # name -> list of (kind, [points...], material)
        self.faces = []
        self.cur = "default"

    def add_face(self, kind, pts, material):
        rec = (kind, pts, material)
        self.groups[self.cur].append(rec)
        self.faces.append(rec)


def load_mtl(path, table):
    """LoadObjMtl: one `Lambertian(Ka)` per `newmtl`, appended to `table` (list of make_materials specs).
    Returns {name: global index}."""
    refs, name, ka = {}, None, (0.0, 0.0, 0.0)

    def flush():
        if name is not None:
            table.append(("lambert", ka))
            refs[name] = len(table) - 1
    with open(path, "r", encoding="utf-8-sig") as fh:
        for line in fh:
            tok = line.split("#", 1)[0].split()
            if not tok:
                continue
            if tok[0] == "newmtl":
                flush()
                name, ka = " ".join(tok[1:]), (0.0, 0.0, 0.0)
            elif tok[0] == "Ka" and len(tok) >= 4:
                ka = tuple(float(x) for x in tok[1:4])
    flush()
    return refs


def load_obj(path, table=None):
    """LoadObjModel (ObjModelLoader.fs:296-340)."""
    table = [] if table is None else table
    st, refs, usemtl, mtl_loaded = ObjState(), {}, "white", False
    base = os.path.dirname(os.path.abspath(path))
    with open(path, "r", encoding="utf-8-sig") as fh:
        lines = fh.read().splitlines()
    # the reference resolves the first `mtllib` of the file before it replays the statements (:311-326)
    for line in lines:
        tok = line.split()
        if tok and tok[0] == "mtllib" and len(tok) > 1:
            refs, mtl_loaded = load_mtl(os.path.join(base, tok[1]), table), True
            break
    for ln, line in enumerate(lines, 1):
        tok = line.split("#", 1)[0].split()
        if not tok:
            continue
        key = tok[0]
        if key == "v":
            if len(tok) != 4:
                raise IngestError(f"{path}:{ln}: a vertex needs 3 coordinates")       # toPoint's assert
            st.vertices.append(tuple(float(x) for x in tok[1:4]))
        elif key == "g":
            st.cur = line.split(None, 1)[1].strip() if len(tok) > 1 else ""            # pLine: the rest of the line
            st.groups.setdefault(st.cur, [])
        elif key == "usemtl":
            usemtl = line.split(None, 1)[1].strip()
        elif key == "f":
            n = len(st.vertices)
            idx = []
            for ref in tok[1:]:
                i = int(ref.split("/")[0])
                idx.append(i - 1 if i > 0 else n + i)                                 # VertexReferencing.VI
            if len(idx) not in (3, 4):
                raise IngestError(f"{path}:{ln}: faces must have 3 or 4 vertices, got {len(idx)}")
            if min(idx) < 0 or max(idx) >= n:
                raise IngestError(f"{path}:{ln}: vertex reference out of range")
            st.add_face(TRIANGLE if len(idx) == 3 else RECT, [st.vertices[i] for i in idx], refs.get(usemtl, 0))
        # vt, vn, o, s, usemap, mtllib, maplib, anything else: read and ignored (see module docstring)
    st.mtl_loaded = mtl_loaded
    return st


def prims_of(faces, material=None):
    """IHitable[] of a face list as MfxPrim records; `material` overrides the faces' own (Scene.fs:154-161)."""
    out = make_prims(len(faces))
    for k, (kind, pts, mat) in enumerate(faces):
        out[k]["kind"] = kind
        out[k]["material"] = mat if material is None else material
        out[k]["v"][: 3 * len(pts)] = np.asarray(pts, np.float64).ravel()
    return out


def _args(node):
    return {n.get("name"): n for n in node}


def _f3(node):
    vs = [x.strip() for x in node.get("value").split(",") if x.strip()]
    if len(vs) != 3:
        raise IngestError(f"expected three comma-separated numbers, got {node.get('value')!r}")
    return tuple(float(x) for x in vs)


def init_scene_state(xml_text, base_dir=".", max_depth=3, integrator=PATH_INTEGRATOR):
    """InitSceneState + the parts of `new Scene(state)` that the path needs (Scene.fs:262-313) -> SceneDesc."""
    root = ET.fromstring(xml_text)
    if root.tag != "Scene" or root.get("version") != "0.1":
        raise IngestError("this scene loader only supports <Scene version=\"0.1\">")
    sect = {}
    for n in root:
        if n.tag not in ("Camera", "Models", "Materials", "Shapes", "Light", "Film"):
            raise IngestError(f"unknown scene element <{n.tag}>")
        sect[n.tag] = n
    for need in ("Camera", "Models", "Materials", "Shapes", "Light", "Film"):
        if need not in sect:
            raise IngestError(f"<{need}> is missing")
    # Camera.ToPinhole (:50-68)
    cam = sect["Camera"]
    if cam.get("type") != "pinhole":
        raise IngestError("only pinhole cameras exist")
    a = _args(cam)
    position = _f3(a["position"]) if "position" in a else (0.0, 0.0, 0.0)
    direction = _f3(a["direction"]) if "direction" in a else (0.0, 0.0, 0.0)
    fov = float(a["fov"].get("value")) if "fov" in a else 60.0
    aspect = float(a["aspectratio"].get("value")) if "aspectratio" in a else 1.333
    # Model.ToModels (:93-134): loading a model appends its MTL materials to the global table right away
    table, models = [], {}
    for m in sect["Models"]:
        if m.tag != "Model" or m.get("type") != "obj":
            raise IngestError("only <Model type=\"obj\"> exists")
        fn = _args(m)["filename"].get("value")
        models[m.get("name")] = load_obj(os.path.join(base_dir, fn), table)

    def find(ref):                                                   # Shape.FindModel (:137-141)
        nm = [x.strip() for x in ref.split(".") if x.strip()]
        if len(nm) != 2 or nm[0] not in models or nm[1] not in models[nm[0]].groups:
            raise IngestError(f"unknown object reference {ref!r}")
        return models[nm[0]].groups[nm[1]]
    # Lights.ToAreaLight (:181-197)
    lt = sect["Light"]
    if lt.get("type") != "area":
        raise IngestError("only area lights exist")
    a = _args(lt)
    faces = find(a["shape_ref"].get("value"))
    if not faces or faces[0][0] != RECT:
        raise IngestError("the light's shape must start with a quad")
    p = np.asarray(faces[0][1], np.float64)
    e1, e2 = p[1] - p[0], p[2] - p[0]
    nrm = np.cross(e1, e2)
    nrm = nrm / np.sqrt((nrm * nrm).sum())                           # Triangle ctor normal (Trangle.fs:108-112)
    light = AreaLight(p, nrm, _f3(a["intensity"]))
    # Film.ToFilm (:204-214)
    a = _args(sect["Film"])
    width = int(a["width"].get("value")) if "width" in a else 800
    height = int(a["height"].get("value")) if "height" in a else 800
    # Material.ToMaterials (:70-91), added to the manager AFTER the models' materials (:258-259)
    for m in sect["Materials"]:
        if m.get("type") != "lambert":
            raise IngestError("the XML loader only knows lambert materials")
        am = _args(m)
        table.append(("lambert", _f3(am["albedo"]) if "albedo" in am else (0.0, 0.0, 0.0)))
    # Shape.ToShapes (:163-177)
    parts = []
    for sh in sect["Shapes"]:
        if sh.get("type") != "shapelist":
            raise IngestError("only <Shape type=\"shapelist\"> exists")
        a = _args(sh)
        mat = int(a["material"].get("value")) if "material" in a else 0
        if not 0 <= mat < len(table):
            raise IngestError(f"material index {mat} outside the global table of {len(table)}")
        parts.append(prims_of(find(a["obj_ref"].get("value")), mat))
    prims = np.concatenate(parts) if parts else make_prims(0)
    if len(prims) == 0:
        raise IngestError("the scene has no shapes")
    camera = PinholeCamera(position, direction, fov, aspect)
    return SceneDesc(prims, make_materials(table), light, camera, width, height, max_depth, integrator)
