// front.cpp — .pbrt scene front end: directives -> flattened pt_scene_desc.
//
// C++ host mirror (rustc is not available) of the reference's L5 layer for the PathIntegrator path:
//   pbrt_parse_file                src/core/parser/parser.rs:60-88, 137-381
//   SceneContext (ParseContext)    src/core/api/scene_context/scene_context.rs
//   create_triangle_mesh_shape     src/shapes/triangle.rs:696-868
//   create_*_material              src/materials/{matte,plastic,glass,metal,mirror}.rs
//   create_{diffuse_area,point,distant}_light   src/lights/*.rs
//   create_film / Film::new        src/core/film/film.rs:64-179, 538-571
//   create_perspective_camera      src/cameras/perspective.rs:345-394, core/camera/projective.rs:23-53
//   create_{halton,sobol,random}_sampler, create_path_integrator, create_bvh_accelerator
// The result is the POD scene description of include/pbrt_r3_b200.h; nothing here touches CUDA.
#include "../../../include/pbrt_r3_b200.h"
#include "params.h"
#include "shape_loaders.h"
#include "spectrum_host.h"
#include "xform.h"

#include <dirent.h>
#include <ftw.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>
#include <sstream>

using namespace pth;

void pt_set_error(const std::string& s);  // api_common.cpp
const pt_options& pt_current_options();    // api_common.cpp: PbrtOptions singleton

struct pt_host_scene {
  std::vector<float> spectra;
  std::vector<float> tri_p, tri_n, tri_uv, tri_s;
  std::vector<pt_spot> spots;
  std::vector<pt_envmap> envmaps;
  std::vector<std::vector<float>> env_texels;  // storage behind pt_envmap.texels
  std::vector<uint32_t> tri_flags;
  std::vector<int32_t> tri_material, tri_light;
  std::vector<std::vector<int32_t>> obj_prims;  // object 0 = top level
  std::vector<uint32_t> obj_first;
  std::vector<int32_t> prims;
  std::vector<pt_instance> instances;
  std::vector<pt_material> materials;
  std::vector<pt_light> lights;        // sampled lights (Scene.lights order)
  std::vector<pt_light> emitters_only; // area lights of instanced shapes (emit when hit, never sampled)
  pt_scene_desc desc;
  std::string filename;
  std::vector<std::string> warnings;
};

namespace {

bool read_text_file(const std::string& path, std::string* out);  // plain or gzipped text (defined below)

struct MaterialInstance {
  std::string name;  // plugin name ("" = none)
  ParamSet params;
};

struct GraphicsState {
  bool has_material = true;
  MaterialInstance material{"matte", ParamSet()};  // graphics_state.rs:54-64 default matte
  std::map<std::string, MaterialInstance> named_materials;
  std::map<std::string, float> float_textures;            // constant float textures
  std::map<std::string, std::vector<float>> spec_textures;  // constant spectrum textures (as rgb)
  std::string area_light_name;
  ParamSet area_light_params;
  bool reverse_orientation = false;
};

struct Builder {
  pt_host_scene* hs;
  SpectrumFactory spec;
  std::vector<Xf> ctm_stack;
  Xf ctm = Xf::identity();
  std::map<std::string, Xf> named_cs;
  std::vector<GraphicsState> gs_stack;
  GraphicsState gs;
  bool in_world = false;
  // render options (render_options.rs:61-88)
  std::string filter_name = "box", film_name = "image", sampler_name = "halton", accel_name = "bvh",
              integrator_name = "path", camera_name = "perspective";
  ParamSet filter_params, film_params, sampler_params, accel_params, integrator_params, camera_params;
  Xf camera_to_world = Xf::identity();
  bool have_camera = false;
  std::map<std::string, int> object_ids;
  int current_object = 0;  // 0 = top level
  int pixelsamples_override = -1;
  std::string cur_dir;  // directory of the file being parsed: relative "filename" parameters resolve against it

  Builder(pt_host_scene* h, int spectrum_kind) : hs(h), spec(spectrum_kind) { hs->obj_prims.resize(1); }

  void warn(const std::string& w) { hs->warnings.push_back(w); }

  // ---- parameter list --------------------------------------------------------------------
  ParamSet parse_params(Lexer& lx) {
    ParamSet ps;
    while (lx.peek().kind == TK_STRING) {
      Token decl = lx.next();
      std::istringstream is(decl.text);
      std::string a, b;
      is >> a >> b;
      std::string type, name;
      if (b.empty()) {
        name = a;
        type = wellknown_type(name);
      } else {
        type = a;
        name = b;
      }
      std::vector<Token> vals;
      Token t = lx.next();
      if (t.kind == TK_LBRACK) {
        for (;;) {
          Token v = lx.next();
          if (v.kind == TK_RBRACK) break;
          if (v.kind == TK_EOF) throw ParseError("unterminated parameter list for \"" + decl.text + "\"");
          vals.push_back(v);
        }
      } else if (t.kind == TK_STRING || t.kind == TK_NUMBER || t.kind == TK_IDENT) {
        vals.push_back(t);
      } else {
        throw ParseError("missing value for parameter \"" + decl.text + "\"");
      }
      Param p;
      auto floats = [&]() { for (auto& v : vals) p.f.push_back(parse_f32(v.text)); };
      if (type == "string") { p.cls = PC_STRING; for (auto& v : vals) p.s.push_back(v.text); }
      else if (type == "texture") { p.cls = PC_TEXTURE; for (auto& v : vals) p.s.push_back(v.text); }
      else if (type == "spectrum") {
        // numeric (lambda, value) pairs are kept as floats under PC_SPECTRUM; file names as strings
        p.cls = PC_SPECTRUM;
        for (auto& v : vals) { if (v.kind == TK_NUMBER) p.f.push_back(parse_f32(v.text)); else p.s.push_back(v.text); }
      }
      else if (type == "bool") {
        p.cls = PC_BOOL;
        for (auto& v : vals) {
          std::string s = v.text;
          for (auto& c : s) c = (char)tolower(c);
          if (s == "true") p.b.push_back(1);
          else if (s == "false") p.b.push_back(0);
          else throw ParseError("bad bool '" + v.text + "'");
        }
      }
      else if (type == "integer") { p.cls = PC_INT; for (auto& v : vals) p.i.push_back((int32_t)strtol(v.text.c_str(), nullptr, 10)); }
      else if (type == "color" || type == "rgb") { p.cls = PC_RGB; floats(); }
      else if (type == "xyz") { p.cls = PC_XYZ; floats(); }
      else if (type == "blackbody") { p.cls = PC_BLACKBODY; floats(); }
      else if (type == "point" || type == "point2" || type == "point3" || type == "point4" || type == "vector" ||
               type == "vector2" || type == "vector3" || type == "vector4" || type == "normal") { p.cls = PC_POINT; floats(); }
      else { p.cls = PC_FLOAT; floats(); }
      ps.set(p.cls, name, p);
    }
    return ps;
  }

  // ---- spectra -----------------------------------------------------------------------------
  int add_spectrum(const std::vector<float>& c) {
    size_t nc = spec.n_channels();
    size_t n = hs->spectra.size() / nc;
    for (size_t k = 0; k < n; k++)
      if (memcmp(&hs->spectra[k * nc], c.data(), nc * sizeof(float)) == 0) return (int)k;
    hs->spectra.insert(hs->spectra.end(), c.begin(), c.end());
    return (int)n;
  }
  // ParamSet::find_one_spectrum / TextureParams::get_spectrum_from (param_set.rs:386-388): rgb -> Reflectance
  bool spectrum_from(const ParamSet& ps, const std::string& key, std::vector<float>* out) {
    if (const Param* q = ps.find(PC_RGB, key)) {
      if (q->f.size() >= 3) { *out = spec.from_rgb(q->f.data(), /*illuminant=*/false); return true; }
    }
    if (const Param* q = ps.find(PC_XYZ, key)) {
      if (q->f.size() >= 3) {
        float rgb[3];
        SpectrumFactory::xyz_to_rgb(q->f.data(), rgb);
        *out = spec.from_rgb(rgb, false);
        return true;
      }
    }
    if (const Param* q = ps.find(PC_SPECTRUM, key)) {
      if (q->f.size() >= 2 && q->f.size() % 2 == 0) {
        std::vector<float> lam, val;
        for (size_t i = 0; i < q->f.size(); i += 2) { lam.push_back(q->f[i]); val.push_back(q->f[i + 1]); }
        *out = spec.from_sampled(lam, val);
        return true;
      }
      if (!q->s.empty()) {
        // The reference loads the file (an unreadable one is reported) and then files the spectrum under the FILE's name
        // (scene_context.rs:744-758: add_spectrum_no_key(name, ..)), where find_one_spectrum(key) never looks: the parameter
        // keeps its default.  Reproduced: the file is opened for the same diagnostic, its values are not used.
        for (const std::string& fn : q->s) {
          std::string path = (fn.size() && fn[0] == '/') ? fn : (cur_dir.empty() ? fn : cur_dir + "/" + fn);
          std::ifstream f(path);
          if (!f) warn("Unable to read SPD file \"" + path + "\".");
        }
        warn("spectrum file parameter \"" + key + "\": pbrt-r3 stores it under the file's name, so the parameter keeps its default (reproduced)");
        return false;
      }
    }
    if (const Param* q = ps.find(PC_BLACKBODY, key)) {  // scene_context.rs:763-778
      if (q->f.size() >= 2) { *out = spec.from_blackbody(q->f); return true; }
    }
    return false;
  }
  std::vector<float> one_spectrum(const ParamSet& ps, const std::string& key, float dflt) {
    std::vector<float> r;
    if (spectrum_from(ps, key, &r)) return r;
    return spec.constant(dflt);
  }

  // ---- TextureParams (texture_params.rs:60-190) for constant textures -------------------------
  struct TP {
    const ParamSet* geom;
    const ParamSet* mat;
  };
  static const std::string* tex_name(const TP& tp, const std::string& key) {
    for (const ParamSet* ps : {tp.geom, tp.mat})
      if (const Param* q = ps->find(PC_TEXTURE, key))
        if (!q->s.empty()) return &q->s[0];
    return nullptr;
  }
  std::vector<float> tp_spectrum(const TP& tp, const std::string& key, const std::vector<float>& dflt) {
    std::vector<float> r;
    if (const std::string* n = tex_name(tp, key)) {
      auto it = gs.spec_textures.find(*n);
      if (it != gs.spec_textures.end()) return it->second;
      if (spectrum_from(*tp.mat, *n, &r) || spectrum_from(*tp.geom, *n, &r)) return r;
      return dflt;
    }
    if (spectrum_from(*tp.mat, key, &r) || spectrum_from(*tp.geom, key, &r)) return r;  // material params FIRST
    return dflt;
  }
  bool tp_float_or_null(const TP& tp, const std::string& key, float* out) {
    if (const std::string* n = tex_name(tp, key)) {
      auto it = gs.float_textures.find(*n);
      if (it != gs.float_textures.end()) { *out = it->second; return true; }
      return false;
    }
    for (const ParamSet* ps : {tp.geom, tp.mat})
      if (const Param* q = ps->find(PC_FLOAT, key))
        if (q->f.size() == 1) { *out = q->f[0]; return true; }
    return false;
  }
  float tp_float(const TP& tp, const std::string& key, float dflt) {
    float v;
    return tp_float_or_null(tp, key, &v) ? v : dflt;
  }
  static bool tp_bool(const TP& tp, const std::string& key, bool dflt) {
    for (const ParamSet* ps : {tp.geom, tp.mat})
      if (const Param* q = ps->find(PC_BOOL, key))
        if (q->b.size() == 1) return q->b[0] != 0;
    return dflt;
  }

  // ---- materials (materials/*.rs create_* + compute_scattering_functions inputs) ---------------
  int make_material(const std::string& name, const ParamSet& geom, const ParamSet& mat) {
    TP tp{&geom, &mat};
    pt_material m;
    memset(&m, 0, sizeof(m));
    m.kd = m.ks = m.kr = m.kt = m.opacity = -1;
    m.remap_roughness = 1;
    std::string nm = name;
    if (nm == "" || nm == "none") return -1;
    if (nm == "mix") {  // scene_context.rs:446-451 + mix.rs:96-103: two named materials (matte from these parameters when unknown)
      if (spec.n_channels() != 3) throw ParseError("Material \"mix\" is only supported with RGB spectra on the GPU path");
      static const ParamSet empty;
      int sub[2];
      const char* keys[2] = {"namedmaterial1", "namedmaterial2"};
      for (int k = 0; k < 2; k++) {
        std::string key;
        for (const ParamSet* ps : {tp.geom, tp.mat})
          if (key.empty()) if (const Param* q = ps->find(PC_STRING, keys[k])) if (q->s.size() == 1) key = q->s[0];
        auto it = gs.named_materials.find(key);
        if (it == gs.named_materials.end()) sub[k] = make_material("matte", geom, mat);
        else {
          if (it->second.name == "mix") throw ParseError("Material \"mix\" of a \"mix\" material is not supported on the GPU path");
          sub[k] = make_material(it->second.name, it->second.params, empty);
        }
        if (sub[k] < 0) throw ParseError("Material \"mix\": \"" + key + "\" has no BSDF (material \"none\"), which the reference does not accept either");
      }
      m.type = PT_MAT_MIX;
      m.kd = sub[0]; m.ks = sub[1];
      m.kr = add_spectrum(tp_spectrum(tp, "amount", spec.constant(0.5f)));
      for (size_t k = 0; k < hs->materials.size(); k++)
        if (memcmp(&hs->materials[k], &m, sizeof(m)) == 0) return (int)k;
      hs->materials.push_back(m);
      return (int)hs->materials.size() - 1;
    }
    if (nm != "matte" && nm != "plastic" && nm != "glass" && nm != "metal" && nm != "mirror" && nm != "uber" && nm != "substrate" && nm != "translucent") {
      // create_material.rs:65-70: unknown names fall back to matte with a warning
      warn("Unable to create material \"" + nm + "\" on the GPU path; using \"matte\".");
      nm = "matte";
    }
    if (nm == "matte") {
      m.type = PT_MAT_MATTE;
      m.kd = add_spectrum(tp_spectrum(tp, "Kd", spec.constant(0.5f)));
      m.sigma = tp_float(tp, "sigma", 0.0f);
    } else if (nm == "plastic") {
      m.type = PT_MAT_PLASTIC;
      m.kd = add_spectrum(tp_spectrum(tp, "Kd", spec.constant(0.25f)));
      m.ks = add_spectrum(tp_spectrum(tp, "Ks", spec.constant(0.25f)));
      m.uroughness = m.vroughness = tp_float(tp, "roughness", 0.1f);
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    } else if (nm == "mirror") {
      m.type = PT_MAT_MIRROR;
      m.kr = add_spectrum(tp_spectrum(tp, "Kr", spec.constant(0.9f)));
    } else if (nm == "glass") {
      m.type = PT_MAT_GLASS;
      m.kr = add_spectrum(tp_spectrum(tp, "Kr", spec.constant(1.0f)));
      m.ks = add_spectrum(tp_spectrum(tp, "Kt", spec.constant(1.0f)));
      m.uroughness = tp_float(tp, "uroughness", 0.0f);
      m.vroughness = tp_float(tp, "vroughness", 0.0f);
      float eta;
      if (!tp_float_or_null(tp, "eta", &eta) && !tp_float_or_null(tp, "index", &eta)) eta = 1.5f;
      m.eta = eta;
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    } else if (nm == "uber") {  // uber.rs:142-168
      if (spec.n_channels() != 3) throw ParseError("Material \"uber\" is only supported with RGB spectra on the GPU path");
      m.type = PT_MAT_UBER;
      m.kd = add_spectrum(tp_spectrum(tp, "Kd", spec.constant(0.25f)));
      m.ks = add_spectrum(tp_spectrum(tp, "Ks", spec.constant(0.25f)));
      m.kr = add_spectrum(tp_spectrum(tp, "Kr", spec.constant(0.0f)));
      m.kt = add_spectrum(tp_spectrum(tp, "Kt", spec.constant(0.0f)));
      m.opacity = add_spectrum(tp_spectrum(tp, "opacity", spec.constant(1.0f)));
      float rough = tp_float(tp, "roughness", 0.1f);
      float ur, vr;
      m.uroughness = tp_float_or_null(tp, "uroughness", &ur) ? ur : rough;
      m.vroughness = tp_float_or_null(tp, "vroughness", &vr) ? vr : rough;
      float eta;
      if (!tp_float_or_null(tp, "eta", &eta) && !tp_float_or_null(tp, "index", &eta)) eta = 1.5f;
      m.eta = eta;
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    } else if (nm == "substrate") {  // substrate.rs:62-78
      if (spec.n_channels() != 3) throw ParseError("Material \"substrate\" is only supported with RGB spectra on the GPU path");
      m.type = PT_MAT_SUBSTRATE;
      m.kd = add_spectrum(tp_spectrum(tp, "Kd", spec.constant(0.5f)));
      m.ks = add_spectrum(tp_spectrum(tp, "Ks", spec.constant(0.5f)));
      m.uroughness = tp_float(tp, "uroughness", 0.1f);
      m.vroughness = tp_float(tp, "vroughness", 0.1f);
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    } else if (nm == "translucent") {  // translucent.rs:102-122
      if (spec.n_channels() != 3) throw ParseError("Material \"translucent\" is only supported with RGB spectra on the GPU path");
      m.type = PT_MAT_TRANSLUCENT;
      m.kd = add_spectrum(tp_spectrum(tp, "Kd", spec.constant(0.25f)));
      m.ks = add_spectrum(tp_spectrum(tp, "Ks", spec.constant(0.25f)));
      m.kr = add_spectrum(tp_spectrum(tp, "reflect", spec.constant(0.5f)));
      m.kt = add_spectrum(tp_spectrum(tp, "transmit", spec.constant(0.5f)));
      m.uroughness = m.vroughness = tp_float(tp, "roughness", 0.1f);
      m.eta = 1.5f;
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    } else {  // metal
      m.type = PT_MAT_METAL;
      m.kr = add_spectrum(tp_spectrum(tp, "eta", spec.copper_n()));
      m.ks = add_spectrum(tp_spectrum(tp, "k", spec.copper_k()));
      float rough = tp_float(tp, "roughness", 0.01f);
      float ur, vr;
      m.uroughness = tp_float_or_null(tp, "uroughness", &ur) ? ur : rough;
      m.vroughness = tp_float_or_null(tp, "vroughness", &vr) ? vr : rough;
      m.remap_roughness = tp_bool(tp, "remaproughness", true);
    }
    if (tex_name(tp, "bumpmap")) throw ParseError("bump maps are not supported on the GPU path");
    for (size_t k = 0; k < hs->materials.size(); k++)
      if (memcmp(&hs->materials[k], &m, sizeof(m)) == 0) return (int)k;
    hs->materials.push_back(m);
    return (int)hs->materials.size() - 1;
  }

  // scene_context.rs:224-267
  static bool shape_may_set_material_parameters(const ParamSet& ps) {
    for (auto& kv : ps.p) {
      const Param& q = kv.second;
      const std::string name = kv.first.substr(kv.first.find(':') + 1);
      switch (q.cls) {
        case PC_TEXTURE: if (name != "alpha" && name != "shadowalpha") return true; break;
        case PC_FLOAT: if (q.f.size() == 1 && name != "radius") return true; break;
        case PC_STRING: if (q.s.size() == 1 && name != "filename" && name != "type" && name != "scheme") return true; break;
        case PC_BOOL: if (q.b.size() == 1) return true; break;
        case PC_INT: if (q.i.size() == 1) return true; break;
        case PC_RGB: case PC_XYZ: case PC_BLACKBODY: case PC_SPECTRUM: if (q.len() >= 1) return true; break;
        case PC_POINT: if (q.f.size() == 3) return true; break;
      }
    }
    return false;
  }
  int material_for_shape(const ParamSet& shape_params) {
    if (!gs.has_material) return -1;
    static const ParamSet empty;
    if (shape_may_set_material_parameters(shape_params)) return make_material(gs.material.name, shape_params, gs.material.params);
    // pbrt_material built it once with (geom = material params, mat = empty) (scene_context.rs:1116-1131)
    return make_material(gs.material.name, gs.material.params, empty);
  }

  // ---- shapes (shapes/triangle.rs:696-868) -----------------------------------------------------
  static bool is_fillable_uv(const std::vector<int32_t>& vi, size_t nverts) {
    std::vector<size_t> chk(nverts, 0);
    for (size_t i = 0; i + 2 < vi.size(); i += 3)
      for (size_t j = 0; j < 3; j++) {
        size_t v = (size_t)(uint32_t)vi[i + j];
        if (v >= nverts) return false;
        if (chk[v] == 0 || chk[v] == j) chk[v] = j;
        else return false;
      }
    return true;
  }

  // raw mesh arrays in object space, before TriangleMesh::new transforms them
  struct MeshData {
    std::vector<int32_t> vi;
    std::vector<float> P, N, S, UV;   // 3, 3, 3, 2 floats per vertex; N / S / UV may be empty
  };

  // PLY reader (shapes/plymesh.rs:253-378 on top of the ply-rs crate): ascii, binary_little_endian and binary_big_endian;
  // element "vertex" with float properties x y z [nx ny nz] [u v | s t | texture_u texture_v | texture_s texture_t],
  // element "face" with an integer list vertex_indices / vertex_index; triangles, and quads split as (0,1,2) (3,0,2).
  // Other properties are skipped; vertex properties the reference panics on (non-float) are an error.
  static std::string load_ply(const std::string& path, MeshData* M) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "cannot open file";
    std::string line;
    if (!std::getline(f, line) || line.substr(0, 3) != "ply") return "not a PLY file";
    enum { ASCII, LE, BE } fmt = ASCII;
    struct Prop { std::string name, type, ctype; bool list; };
    struct Elem { std::string name; size_t count; std::vector<Prop> props; };
    std::vector<Elem> elems;
    bool header_done = false;
    while (std::getline(f, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      std::istringstream is(line);
      std::string w;
      is >> w;
      if (w == "format") { std::string t; is >> t; fmt = t == "ascii" ? ASCII : (t == "binary_little_endian" ? LE : BE); }
      else if (w == "element") { Elem e; is >> e.name >> e.count; elems.push_back(e); }
      else if (w == "property") {
        if (elems.empty()) return "property before element";
        Prop p; std::string t; is >> t;
        if (t == "list") { p.list = true; is >> p.ctype >> p.type >> p.name; }
        else { p.list = false; p.type = t; is >> p.name; }
        elems.back().props.push_back(p);
      } else if (w == "end_header") { header_done = true; break; }
    }
    if (!header_done) return "missing end_header";
    auto tsize = [](const std::string& t) -> int {
      if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
      if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
      if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
      if (t == "double" || t == "float64") return 8;
      return 0;
    };
    auto is_float = [](const std::string& t) { return t == "float" || t == "float32"; };
    // one scalar of type t as double
    auto read_scalar = [&](const std::string& t, double* out) -> bool {
      if (fmt == ASCII) { return (bool)(f >> *out); }
      int n = tsize(t);
      unsigned char b[8];
      if (n == 0 || !f.read((char*)b, n)) return false;
      if (fmt == BE) std::reverse(b, b + n);
      if (t == "char" || t == "int8") *out = (double)(int8_t)b[0];
      else if (t == "uchar" || t == "uint8") *out = (double)b[0];
      else if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); *out = v; }
      else if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); *out = v; }
      else if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); *out = v; }
      else if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); *out = v; }
      else if (is_float(t)) { float v; memcpy(&v, b, 4); *out = v; }
      else { double v; memcpy(&v, b, 8); *out = v; }
      return true;
    };
    for (const Elem& e : elems) {
      if (e.name == "vertex") {
        int ix[8] = {-1, -1, -1, -1, -1, -1, -1, -1};  // x y z nx ny nz u v
        for (size_t k = 0; k < e.props.size(); k++) {
          const std::string& n = e.props[k].name;
          int slot = n == "x" ? 0 : n == "y" ? 1 : n == "z" ? 2 : n == "nx" ? 3 : n == "ny" ? 4 : n == "nz" ? 5
                     : (n == "u" || n == "s" || n == "texture_u" || n == "texture_s") ? 6 : (n == "v" || n == "t" || n == "texture_v" || n == "texture_t") ? 7 : -1;
          if (e.props[k].list) return "list property in the vertex element";
          if (slot >= 0 && !is_float(e.props[k].type)) return "vertex property \"" + n + "\" must be float";
          if (slot >= 0) ix[slot] = (int)k;
        }
        bool has_p = ix[0] >= 0 || ix[1] >= 0 || ix[2] >= 0, has_n = ix[3] >= 0 || ix[4] >= 0 || ix[5] >= 0, has_uv = ix[6] >= 0 || ix[7] >= 0;
        std::vector<double> row(e.props.size());
        for (size_t i = 0; i < e.count; i++) {
          for (size_t k = 0; k < e.props.size(); k++) if (!read_scalar(e.props[k].type, &row[k])) return "truncated vertex data";
          auto val = [&](int s) { return ix[s] >= 0 ? (float)row[(size_t)ix[s]] : 0.0f; };
          if (has_p) { M->P.push_back(val(0)); M->P.push_back(val(1)); M->P.push_back(val(2)); }
          if (has_n) { M->N.push_back(val(3)); M->N.push_back(val(4)); M->N.push_back(val(5)); }
          if (has_uv) { M->UV.push_back(val(6)); M->UV.push_back(val(7)); }
        }
      } else {
        const bool face = e.name == "face";
        for (size_t i = 0; i < e.count; i++) {
          for (const Prop& p : e.props) {
            if (!p.list) { double d; if (!read_scalar(p.type, &d)) return "truncated data"; continue; }
            double cnt;
            if (!read_scalar(p.ctype, &cnt)) return "truncated list";
            std::vector<int32_t> idx((size_t)cnt);
            for (auto& v : idx) { double d; if (!read_scalar(p.type, &d)) return "truncated list"; v = (int32_t)d; }
            if (face && (p.name == "vertex_indices" || p.name == "vertex_index")) {
              if (idx.size() == 3) M->vi.insert(M->vi.end(), idx.begin(), idx.end());
              else if (idx.size() == 4) { const int32_t q[6] = {idx[0], idx[1], idx[2], idx[3], idx[0], idx[2]}; M->vi.insert(M->vi.end(), q, q + 6); }
              else return "face with " + std::to_string(idx.size()) + " vertices (only triangles and quads are supported)";
            }
          }
        }
      }
    }
    return "";
  }

  void shape(const std::string& name, const ParamSet& ps_in) {
    if (name != "trianglemesh" && name != "plymesh" && name != "heightfield" && name != "loopsubdiv" && name != "nurbs")
      throw ParseError("Shape \"" + name + "\" is not supported on the GPU path (triangle meshes: trianglemesh, plymesh, heightfield, loopsubdiv, nurbs)");
    ParamSet ps = ps_in;
    bool is_area = !gs.area_light_name.empty();
    if (is_area) {  // scene_context.rs:405-421
      bool two_sided = gs.area_light_params.one_bool("twosided", true);
      two_sided = ps.one_bool("twosided", two_sided);
      Param b; b.cls = PC_BOOL; b.b.push_back(two_sided ? 1 : 0);
      ps.set(PC_BOOL, "twosided", b);
    }
    if (ps.find(PC_TEXTURE, "alpha") || ps.find(PC_FLOAT, "alpha") || ps.find(PC_TEXTURE, "shadowalpha") || ps.find(PC_FLOAT, "shadowalpha"))
      throw ParseError("alpha-masked shapes are not supported on the GPU path");
    MeshData M;
    if (name == "heightfield" || name == "loopsubdiv" || name == "nurbs") {  // shapes/{heightfield,loopsubdiv,nurbs}.rs: meshes by construction
      pth::LoadedMesh L;
      try {
        if (name == "heightfield") {
          const Param* pz = ps.find(PC_FLOAT, "Pz");
          L = pth::heightfield_mesh(ps.one_int("nu", -1), ps.one_int("nv", -1), pz ? pz->f : std::vector<float>());
        } else if (name == "loopsubdiv") {
          const Param* pi = ps.find(PC_INT, "indices");
          const Param* pp = ps.find(PC_POINT, "P");
          L = pth::loop_subdiv_mesh(ps.one_int("levels", ps.one_int("nlevels", 3)), pi ? pi->i : std::vector<int32_t>(), pp ? pp->f : std::vector<float>());
        } else {  // nurbs.rs:248-404
          const Param* uk = ps.find(PC_FLOAT, "uknots");
          const Param* vk = ps.find(PC_FLOAT, "vknots");
          const Param* pp = ps.find(PC_POINT, "P");
          const Param* pw = (pp && !pp->f.empty()) ? nullptr : ps.find(PC_POINT, "Pw");
          const Param* q = pw ? pw : pp;
          float r[4];
          const float* rp[4];
          const char* rn[4] = {"u0", "u1", "v0", "v1"};
          for (int k = 0; k < 4; k++) { const Param* f = ps.find(PC_FLOAT, rn[k]); rp[k] = (f && f->f.size() == 1) ? &(r[k] = f->f[0]) : nullptr; }
          L = pth::nurbs_mesh(ps.one_int("nu", -1), ps.one_int("uorder", -1), uk ? uk->f : std::vector<float>(), ps.one_int("nv", -1), ps.one_int("vorder", -1),
                              vk ? vk->f : std::vector<float>(), q ? q->f : std::vector<float>(), pw != nullptr, rp[0], rp[1], rp[2], rp[3], ps.one_int("diceu", 30),
                              ps.one_int("dicev", 30));
        }
      } catch (const std::exception& e) { throw ParseError(std::string("Shape \"") + name + "\": " + e.what()); }
      M.vi = L.vi; M.P = L.P; M.N = L.N; M.UV = L.UV;
      // create_heightfield / create_tesselated_mesh hand create_triangle_mesh an EMPTY parameter set (heightfield.rs:62,
      // nurbs.rs:231): always two-sided
      emit_mesh(name == "loopsubdiv" ? ps : ParamSet(), ps_in, is_area, M);
      return;
    }
    if (name == "plymesh") {  // shapes/plymesh.rs:253-378
      std::string fn = ps.one_string("filename", "");
      std::string path = (fn.size() && fn[0] == '/') ? fn : (cur_dir.empty() ? fn : cur_dir + "/" + fn);
      std::string err = load_ply(path, &M);
      if (!err.empty()) throw ParseError("plymesh \"" + path + "\": " + err);
      if (M.vi.empty() || M.P.empty()) { warn("Invalid mesh"); return; }
    } else {
      const Param* pi = ps.find(PC_INT, "indices");
      const Param* pp = ps.find(PC_POINT, "P");
      if (!pi || !pp || pi->i.empty() || pp->f.empty()) { warn("Invalid mesh"); return; }
      M.vi = pi->i;
      M.P.assign(pp->f.begin(), pp->f.begin() + (pp->f.size() / 3) * 3);
      size_t nverts = M.P.size() / 3;
      const Param* puv = ps.find(PC_FLOAT, "uv");
      if (!puv) puv = ps.find(PC_FLOAT, "st");
      if (puv) {
        M.UV.assign(puv->f.begin(), puv->f.begin() + (puv->f.size() / 2) * 2);
      } else if (is_fillable_uv(M.vi, nverts)) {
        const float tri_uv[3][2] = {{0, 0}, {1, 0}, {1, 1}};
        M.UV.assign(nverts * 2, 0.0f);
        for (size_t i = 0; i + 2 < M.vi.size(); i += 3)
          for (int j = 0; j < 3; j++) {
            size_t v = (size_t)M.vi[i + j];
            if (v < nverts && M.UV[2 * v] == 0.0f && M.UV[2 * v + 1] == 0.0f) { M.UV[2 * v] = tri_uv[j][0]; M.UV[2 * v + 1] = tri_uv[j][1]; }
          }
      }
      if (const Param* pst = ps.find(PC_POINT, "S")) M.S.assign(pst->f.begin(), pst->f.begin() + (pst->f.size() / 3) * 3);  // triangle.rs:823-829
      if (const Param* pn = ps.find(PC_POINT, "N")) M.N.assign(pn->f.begin(), pn->f.begin() + (pn->f.size() / 3) * 3);
    }
    emit_mesh(ps, ps_in, is_area, M);
  }

  // TriangleMesh::new + create_triangle_mesh (triangle.rs:37-92, 696-754) for either source of the arrays
  void emit_mesh(const ParamSet& ps, const ParamSet& ps_in, bool is_area, const MeshData& M) {
    const std::vector<int32_t>& vi = M.vi;
    size_t nverts = M.P.size() / 3;
    for (int32_t v : vi) if (v < 0 || (size_t)v >= nverts) throw ParseError("triangle mesh index out of range");
    const Xf& o2w = ctm;
    std::vector<V3> P(nverts), N, Sv;
    const std::vector<float>& UV = M.UV;
    for (size_t i = 0; i < nverts; i++) P[i] = m_point(o2w.m, v3(M.P[3 * i], M.P[3 * i + 1], M.P[3 * i + 2]));
    if (!UV.empty() && UV.size() / 2 < nverts) throw ParseError("triangle mesh: fewer uv values than vertices");
    if (!M.S.empty()) {  // TriangleMesh::new transforms the tangents as vectors (triangle.rs:55-60)
      size_t nsv = M.S.size() / 3;
      if (nsv < nverts) throw ParseError("triangle mesh: fewer tangents than vertices");
      Sv.resize(nsv);
      for (size_t i = 0; i < nsv; i++) Sv[i] = m_vector(o2w.m, v3(M.S[3 * i], M.S[3 * i + 1], M.S[3 * i + 2]));
    }
    if (!M.N.empty()) {
      size_t nn = M.N.size() / 3;
      if (nn < nverts) throw ParseError("triangle mesh: fewer normals than vertices");
      N.resize(nn);
      for (size_t i = 0; i < nn; i++) N[i] = m_normal_with_inv(o2w.minv, v3(M.N[3 * i], M.N[3 * i + 1], M.N[3 * i + 2]));
    }
    bool two_sided = ps.one_bool("twosided", true);
    uint32_t flags = 0;
    if (gs.reverse_orientation) flags |= PT_TRI_REVERSE_ORIENTATION;
    if (x_swaps_handedness(o2w)) flags |= PT_TRI_SWAPS_HANDEDNESS;
    if (two_sided) flags |= PT_TRI_TWO_SIDED;
    if (!N.empty()) flags |= PT_TRI_HAS_N;
    if (!UV.empty()) flags |= PT_TRI_HAS_UV;
    if (!Sv.empty()) flags |= PT_TRI_HAS_S;

    int mat = material_for_shape(ps_in);
    int light_spec = -1, light_two_sided = 0, light_samples = 1;
    if (is_area) {
      if (gs.area_light_name != "diffuse" && gs.area_light_name != "area")
        throw ParseError("AreaLightSource \"" + gs.area_light_name + "\" unknown");
      std::vector<float> L = one_spectrum(gs.area_light_params, "L", 1.0f);
      std::vector<float> sc = one_spectrum(gs.area_light_params, "scale", 1.0f);
      for (size_t k = 0; k < L.size(); k++) L[k] = L[k] * sc[k];
      light_spec = add_spectrum(L);
      light_two_sided = gs.area_light_params.one_bool("twosided", false) ? 1 : 0;
      light_samples = gs.area_light_params.one_int("nsamples", 1);  // lights/diffuse.rs:177-185
      if (pt_current_options().quick_render) light_samples = std::max(light_samples / 4, 1);
      if (light_samples < 1) throw ParseError("AreaLightSource: nsamples must be positive");
    }
    bool want_n = !N.empty(), want_uv = !UV.empty(), want_s = !Sv.empty();
    size_t ntri = vi.size() / 3;
    for (size_t t = 0; t < ntri; t++) {
      size_t i0 = (size_t)vi[3 * t], i1 = (size_t)vi[3 * t + 1], i2 = (size_t)vi[3 * t + 2];
      V3 p0 = P[i0], p1 = P[i1], p2 = P[i2];
      float area = 0.5f * vlen(vcross(p1 - p0, p2 - p0));
      if (!(area > 1e-16f)) continue;  // triangle.rs:726
      int32_t tri = (int32_t)hs->tri_flags.size();
      const V3 pv[3] = {p0, p1, p2};
      for (int k = 0; k < 3; k++) { hs->tri_p.push_back(pv[k].x); hs->tri_p.push_back(pv[k].y); hs->tri_p.push_back(pv[k].z); }
      if (want_n || !hs->tri_n.empty()) {
        hs->tri_n.resize((size_t)tri * 9, 0.0f);
        const size_t ii[3] = {i0, i1, i2};
        for (int k = 0; k < 3; k++) {
          V3 n = want_n ? N[ii[k]] : v3(0, 0, 0);
          hs->tri_n.push_back(n.x); hs->tri_n.push_back(n.y); hs->tri_n.push_back(n.z);
        }
      }
      if (want_s || !hs->tri_s.empty()) {
        hs->tri_s.resize((size_t)tri * 9, 0.0f);
        const size_t ii[3] = {i0, i1, i2};
        for (int k = 0; k < 3; k++) {
          V3 sv = want_s ? Sv[ii[k]] : v3(0, 0, 0);
          hs->tri_s.push_back(sv.x); hs->tri_s.push_back(sv.y); hs->tri_s.push_back(sv.z);
        }
      }
      if (want_uv || !hs->tri_uv.empty()) {
        hs->tri_uv.resize((size_t)tri * 6, 0.0f);
        const size_t ii[3] = {i0, i1, i2};
        for (int k = 0; k < 3; k++) {
          hs->tri_uv.push_back(want_uv ? UV[2 * ii[k]] : 0.0f);
          hs->tri_uv.push_back(want_uv ? UV[2 * ii[k] + 1] : 0.0f);
        }
      }
      hs->tri_flags.push_back(flags);
      hs->tri_material.push_back(mat);
      int32_t li = -1;
      if (is_area) {
        pt_light l;
        memset(&l, 0, sizeof(l));
        l.type = PT_LIGHT_AREA;
        l.spectrum = light_spec;
        l.tri = tri;
        l.two_sided = light_two_sided;
        l.n_samples = light_samples;
        if (current_object == 0) {
          hs->lights.push_back(l);
          li = (int32_t)hs->lights.size() - 1;
        } else {  // "Area lights not supported with object instancing": still emits when hit
          hs->emitters_only.push_back(l);
          li = -2 - ((int32_t)hs->emitters_only.size() - 1);  // fixed up in finish()
        }
      }
      hs->tri_light.push_back(li);
      hs->obj_prims[(size_t)current_object].push_back(tri);
    }
  }

  // ---- lights (lights/{point,distant}.rs create_*) ---------------------------------------------
  void light_source(const std::string& name, const ParamSet& ps) {
    pt_light l;
    memset(&l, 0, sizeof(l));
    l.tri = -1;
    l.n_samples = 1;  // Light::get_sample_count default (core/light/light.rs:76-78)
    if (name == "point") {
      std::vector<float> I = one_spectrum(ps, "I", 1.0f), sc = one_spectrum(ps, "scale", 1.0f);
      for (size_t k = 0; k < I.size(); k++) I[k] = I[k] * sc[k];
      float from[3] = {0, 0, 0};
      ps.one_point("from", from);
      Xf l2w = x_translate(from[0], from[1], from[2]) * ctm;  // order swapped vs v3 (lights/point.rs:121)
      V3 p = m_point(l2w.m, v3(0, 0, 0));
      l.type = PT_LIGHT_POINT;
      l.spectrum = add_spectrum(I);
      l.v[0] = p.x; l.v[1] = p.y; l.v[2] = p.z;
    } else if (name == "distant") {
      std::vector<float> L = one_spectrum(ps, "L", 1.0f), sc = one_spectrum(ps, "scale", 1.0f);
      for (size_t k = 0; k < L.size(); k++) L[k] = L[k] * sc[k];
      float from[3] = {0, 0, 0}, to[3] = {0, 0, 1};
      ps.one_point("from", from);
      ps.one_point("to", to);
      V3 dir = v3(from[0], from[1], from[2]) - v3(to[0], to[1], to[2]);
      V3 w = vnorm(m_vector(ctm.m, dir));
      l.type = PT_LIGHT_DISTANT;
      l.spectrum = add_spectrum(L);
      l.v[0] = w.x; l.v[1] = w.y; l.v[2] = w.z;
    } else if (name == "spot") {  // lights/spot.rs:136-166
      std::vector<float> I = one_spectrum(ps, "I", 1.0f), sc = one_spectrum(ps, "scale", 1.0f);
      for (size_t k = 0; k < I.size(); k++) I[k] = I[k] * sc[k];
      float coneangle = ps.one_float("coneangle", 30.0f);
      float conedelta = ps.one_float("conedelta", 5.0f);
      conedelta = ps.one_float("conedeltaangle", conedelta);
      float from[3] = {0, 0, 0}, to[3] = {0, 0, 1};
      ps.one_point("from", from);
      ps.one_point("to", to);
      V3 dir = vnorm(v3(to[0], to[1], to[2]) - v3(from[0], from[1], from[2]));
      // Vector3f::coordinate_system (vector3.rs:72-83)
      V3 v1 = vnorm(dir), v2;
      if (fabsf(v1.x) > fabsf(v1.y)) v2 = vnorm(v3(-v1.z, 0.0f, v1.x));
      else v2 = vnorm(v3(0.0f, v1.z, -v1.y));
      V3 v3_ = vnorm(vcross(v1, v2));
      M4 dz = M4::identity();
      dz.m[0] = v2.x; dz.m[1] = v2.y; dz.m[2] = v2.z;
      dz.m[4] = v3_.x; dz.m[5] = v3_.y; dz.m[6] = v3_.z;
      dz.m[8] = v1.x; dz.m[9] = v1.y; dz.m[10] = v1.z;
      Xf l2w = ctm * x_translate(from[0], from[1], from[2]) * xinverse(x_from_matrix(dz));
      V3 p = m_point(l2w.m, v3(0, 0, 0));
      pt_spot sp;
      memset(&sp, 0, sizeof(sp));
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) sp.w2l[3 * r + c] = l2w.minv.m[4 * r + c];
      sp.cos_total_width = cosf(radians(coneangle));
      sp.cos_falloff_start = cosf(radians(coneangle - conedelta));
      l.type = PT_LIGHT_SPOT;
      l.spectrum = add_spectrum(I);
      l.tri = (int32_t)hs->spots.size();
      l.v[0] = p.x; l.v[1] = p.y; l.v[2] = p.z;
      hs->spots.push_back(sp);
    } else if (name == "infinite" || name == "exinfinite") {  // lights/infinite.rs:271-295 (create_light.rs: both names)
      if (spec.n_channels() != 3) throw ParseError("LightSource \"infinite\" is only supported with RGB spectra on the GPU path");
      std::vector<float> L = one_spectrum(ps, "L", 1.0f), sc = one_spectrum(ps, "scale", 1.0f);
      for (size_t k = 0; k < L.size(); k++) L[k] = L[k] * sc[k];
      std::string fn = ps.one_string("mapname", "");
      l.n_samples = ps.one_int("samples", ps.one_int("nsamples", 1));
      if (pt_current_options().quick_render) l.n_samples = std::max(l.n_samples / 4, 1);
      pt_envmap em;
      memset(&em, 0, sizeof(em));
      std::vector<float> tex;
      if (fn.empty()) { em.width = em.height = 1; tex = {L[0], L[1], L[2]}; }
      else {  // make_mipmap (infinite.rs:47-58): texels clamped at zero, times L * scale
        std::string path = (fn[0] == '/') ? fn : (cur_dir.empty() ? fn : cur_dir + "/" + fn);
        std::string err = load_pfm(path, &tex, &em.width, &em.height);
        if (!err.empty()) throw ParseError("LightSource \"infinite\" mapname \"" + path + "\": " + err);
        if (em.width > (1 << 14) || em.height > (1 << 14)) throw ParseError("LightSource \"infinite\": environment maps above 16384 pixels a side are not supported");
        for (size_t i = 0; i < tex.size(); i++) tex[i] = std::max(tex[i], 0.0f) * L[i % 3];
      }
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { em.l2w[3 * r + c] = ctm.m.m[4 * r + c]; em.w2l[3 * r + c] = ctm.minv.m[4 * r + c]; }
      l.type = PT_LIGHT_INFINITE;
      l.spectrum = add_spectrum(L);
      l.tri = (int32_t)hs->envmaps.size();
      hs->envmaps.push_back(em);
      hs->env_texels.push_back(std::move(tex));
    } else if (name == "projection" || name == "goniometric") {  // lights/projection.rs:176-195, goniometric.rs:143-159
      if (spec.n_channels() != 3) throw ParseError("LightSource \"" + name + "\" is only supported with RGB spectra on the GPU path");
      const bool proj = name == "projection";
      std::vector<float> I = one_spectrum(ps, proj ? "I" : "L", 1.0f), sc = one_spectrum(ps, "scale", 1.0f);
      for (size_t k = 0; k < I.size(); k++) I[k] = I[k] * sc[k];
      std::string fn = ps.one_string("mapname", "");
      pt_envmap em;
      memset(&em, 0, sizeof(em));
      std::vector<float> tex;
      if (fn.empty()) { em.width = em.height = 1; tex = {1.0f, 1.0f, 1.0f}; }
      else {  // make_mipmap: texels clamped at zero
        std::string path = (fn[0] == '/') ? fn : (cur_dir.empty() ? fn : cur_dir + "/" + fn);
        std::string err = load_pfm(path, &tex, &em.width, &em.height);
        if (!err.empty()) throw ParseError("LightSource \"" + name + "\" mapname \"" + path + "\": " + err);
        if (em.width > (1 << 14) || em.height > (1 << 14)) throw ParseError("LightSource \"" + name + "\": maps above 16384 pixels a side are not supported");
        for (size_t i = 0; i < tex.size(); i++) tex[i] = std::max(tex[i], 0.0f);
      }
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { em.l2w[3 * r + c] = ctm.m.m[4 * r + c]; em.w2l[3 * r + c] = ctm.minv.m[4 * r + c]; }
      if (proj) {  // ProjectionLight::new (projection.rs:21-62)
        const float fov = ps.one_float("fov", 45.0f);
        const float aspect = (float)em.width / (float)em.height;
        if (aspect > 1.0f) { em.screen_bounds[0] = -aspect; em.screen_bounds[1] = -1.0f; em.screen_bounds[2] = aspect; em.screen_bounds[3] = 1.0f; }
        else { em.screen_bounds[0] = -1.0f; em.screen_bounds[1] = -1.0f / aspect; em.screen_bounds[2] = 1.0f; em.screen_bounds[3] = 1.0f / aspect; }
        em.hitcher = 1e-3f;
        Xf light_projection = x_perspective(fov, 1e-3f, 1e30f);
        memcpy(em.xf, light_projection.m.m, sizeof(em.xf));
        V3 w_corner = vnorm(m_point(light_projection.minv, v3(em.screen_bounds[2], em.screen_bounds[3], 0.0f)));
        em.cos_total_width = w_corner.z;
      } else {
        memcpy(em.xf, ctm.minv.m, sizeof(em.xf));
      }
      V3 p = m_point(ctm.m, v3(0, 0, 0));
      l.type = proj ? PT_LIGHT_PROJECTION : PT_LIGHT_GONIOMETRIC;
      l.spectrum = add_spectrum(I);
      l.v[0] = p.x; l.v[1] = p.y; l.v[2] = p.z;
      l.tri = (int32_t)hs->envmaps.size();
      hs->envmaps.push_back(em);
      hs->env_texels.push_back(std::move(tex));
    } else {
      throw ParseError("LightSource \"" + name + "\" is not supported on the GPU path (point, spot, distant, infinite, projection, goniometric, area lights only)");
    }
    hs->lights.push_back(l);
  }

  // read_image for ".pfm" (core/imageio/read_image_pfm.rs:25-106): "PF" / "Pf", width, height, scale (negative = little
  // endian), rows bottom to top in the file -> row 0 = top here, values times |scale|, grey replicated to RGB
  static std::string load_pfm(const std::string& path, std::vector<float>* rgb, int32_t* w, int32_t* h) {
    if (path.size() < 4 || path.substr(path.size() - 4) != ".pfm") return "only .pfm maps are supported on the GPU path";
    std::ifstream f(path, std::ios::binary);
    if (!f) return "cannot open file";
    std::string magic;
    long long width = 0, height = 0;
    float scale = 0.0f;
    f >> magic >> width >> height >> scale;
    if (!f || (magic != "PF" && magic != "Pf") || width <= 0 || height <= 0 || width > 65536 || height > 65536) return "bad PFM header";
    f.get();  // the single whitespace after the scale
    const int nc = magic == "PF" ? 3 : 1;
    const bool le = scale < 0.0f;
    std::vector<unsigned char> raw((size_t)width * height * nc * 4);
    f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)raw.size());
    if ((size_t)f.gcount() != raw.size()) return "truncated PFM data";
    rgb->assign((size_t)width * height * 3, 0.0f);
    const float a = fabsf(scale);
    size_t k = 0;
    for (long long y = 0; y < height; y++) {
      const long long yy = height - y - 1;
      for (long long x = 0; x < width; x++)
        for (int c = 0; c < nc; c++, k++) {
          const unsigned char* q = raw.data() + 4 * k;
          uint32_t bits = le ? ((uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24)
                             : ((uint32_t)q[3] | (uint32_t)q[2] << 8 | (uint32_t)q[1] << 16 | (uint32_t)q[0] << 24);
          float v;
          memcpy(&v, &bits, 4);
          if (a != 1.0f) v *= a;
          float* dst = rgb->data() + 3 * ((size_t)yy * width + x);
          if (nc == 3) dst[c] = v; else dst[0] = dst[1] = dst[2] = v;
        }
    }
    *w = (int32_t)width; *h = (int32_t)height;
    return "";
  }

  // ---- directives -----------------------------------------------------------------------------
  static void read_floats(Lexer& lx, int n, float* out) {
    for (int i = 0; i < n; i++) {
      Token t = lx.next();
      if (t.kind != TK_NUMBER) throw ParseError("expected a number, got '" + t.text + "'");
      out[i] = parse_f32(t.text);
    }
  }
  static std::vector<float> read_float_list(Lexer& lx) {
    std::vector<float> v;
    Token t = lx.next();
    if (t.kind != TK_LBRACK) throw ParseError("expected '['");
    for (;;) {
      t = lx.next();
      if (t.kind == TK_RBRACK) break;
      if (t.kind != TK_NUMBER) throw ParseError("expected a number in list");
      v.push_back(parse_f32(t.text));
    }
    return v;
  }
  static std::string read_string(Lexer& lx) {
    Token t = lx.next();
    if (t.kind != TK_STRING) throw ParseError("expected a string, got '" + t.text + "'");
    return t.text;
  }

  void parse_text(const std::string& text, const std::string& dir) {
    struct DirGuard {
      std::string& ref; std::string saved;
      DirGuard(std::string& r, const std::string& v) : ref(r), saved(r) { r = v; }
      ~DirGuard() { ref = saved; }
    } dir_guard(cur_dir, dir);
    Lexer lx(text);
    for (;;) {
      Token t = lx.next();
      if (t.kind == TK_EOF) break;
      if (t.kind != TK_IDENT) throw ParseError("unexpected token '" + t.text + "'");
      const std::string& d = t.text;
      if (d == "Identity") ctm = Xf::identity();
      else if (d == "Translate") { float v[3]; read_floats(lx, 3, v); ctm = ctm * x_translate(v[0], v[1], v[2]); }
      else if (d == "Scale") { float v[3]; read_floats(lx, 3, v); ctm = ctm * x_scale(v[0], v[1], v[2]); }
      else if (d == "Rotate") { float v[4]; read_floats(lx, 4, v); ctm = ctm * x_rotate(v[0], v[1], v[2], v[3]); }
      else if (d == "LookAt") {
        float v[9]; read_floats(lx, 9, v);
        Xf t2;
        if (!x_look_at(v, v + 3, v + 6, &t2)) throw ParseError("LookAt: degenerate up vector");
        ctm = ctm * t2;
      }
      else if (d == "Transform" || d == "ConcatTransform") {
        std::vector<float> v = read_float_list(lx);
        if (v.size() != 16) throw ParseError(d + " needs 16 values");
        M4 m;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m.m[4 * r + c] = v[(size_t)(4 * c + r)];  // column-major on disk
        M4 im;
        if (!minverse(m, &im)) { warn("Singular matrix in MatrixInvert"); continue; }
        Xf x{m, im};
        ctm = (d == "Transform") ? x : ctm * x;
      }
      else if (d == "CoordinateSystem") named_cs[read_string(lx)] = ctm;
      else if (d == "CoordSysTransform") {
        std::string n = read_string(lx);
        auto it = named_cs.find(n);
        if (it != named_cs.end()) ctm = it->second; else warn("Couldn't find named coordinate system \"" + n + "\"");
      }
      else if (d == "ActiveTransform") { lx.next(); }
      else if (d == "TransformTimes") { float v[2]; read_floats(lx, 2, v); }
      else if (d == "Camera") {
        camera_name = read_string(lx);
        camera_params = parse_params(lx);
        camera_to_world = xinverse(ctm);
        named_cs["camera"] = camera_to_world;
        have_camera = true;
      }
      else if (d == "Film") { film_name = read_string(lx); film_params = parse_params(lx); }
      else if (d == "Sampler") { sampler_name = read_string(lx); sampler_params = parse_params(lx); }
      else if (d == "PixelFilter") { filter_name = read_string(lx); filter_params = parse_params(lx); }
      else if (d == "Integrator" || d == "SurfaceIntegrator") { integrator_name = read_string(lx); integrator_params = parse_params(lx); }
      else if (d == "VolumeIntegrator") { read_string(lx); parse_params(lx); }
      else if (d == "Accelerator") { accel_name = read_string(lx); accel_params = parse_params(lx); }
      else if (d == "WorldBegin") { in_world = true; ctm = Xf::identity(); named_cs["world"] = ctm; }
      else if (d == "WorldEnd") { /* nothing */ }
      else if (d == "AttributeBegin") { gs_stack.push_back(gs); ctm_stack.push_back(ctm); }
      else if (d == "AttributeEnd") {
        if (gs_stack.empty()) throw ParseError("Unmatched AttributeEnd");
        gs = gs_stack.back(); gs_stack.pop_back();
        ctm = ctm_stack.back(); ctm_stack.pop_back();
      }
      else if (d == "TransformBegin") ctm_stack.push_back(ctm);
      else if (d == "TransformEnd") {
        if (ctm_stack.empty()) throw ParseError("Unmatched TransformEnd");
        ctm = ctm_stack.back(); ctm_stack.pop_back();
      }
      else if (d == "ReverseOrientation") gs.reverse_orientation = !gs.reverse_orientation;
      else if (d == "Material") {
        std::string n = read_string(lx);
        ParamSet ps = parse_params(lx);
        if (n == "") gs.has_material = false;
        else { gs.has_material = true; gs.material = MaterialInstance{n, ps}; }
      }
      else if (d == "MakeNamedMaterial") {
        std::string n = read_string(lx);
        ParamSet ps = parse_params(lx);
        std::string type = ps.one_string("type", "");
        if (type.empty()) warn("No parameter string \"type\" found in MakeNamedMaterial");
        gs.named_materials[n] = MaterialInstance{type, ps};
      }
      else if (d == "NamedMaterial") {
        std::string n = read_string(lx);
        auto it = gs.named_materials.find(n);
        if (it == gs.named_materials.end()) warn("NamedMaterial \"" + n + "\" unknown.");
        else { gs.has_material = true; gs.material = it->second; }
      }
      else if (d == "Texture") {
        std::string n = read_string(lx), type = read_string(lx), cls = read_string(lx);
        ParamSet ps = parse_params(lx);
        // Textures whose value does not depend on the surface point fold to constants on the host: "constant", and
        // "scale" / "mix" of such textures (textures/{constant,scale,mix}.rs; same f32 arithmetic as their evaluate()).
        if (cls != "constant" && cls != "scale" && cls != "mix")
          throw ParseError("Texture class \"" + cls + "\" is not supported on the GPU path (constant, scale, mix of constants)");
        if (type != "float" && type != "spectrum" && type != "color") throw ParseError("Texture type \"" + type + "\" unknown");
        static const ParamSet empty_mat;
        TP tp{&ps, &empty_mat};
        const bool is_float = type == "float";
        if (cls == "constant") {
          if (is_float) gs.float_textures[n] = ps.one_float("value", 1.0f);
          else gs.spec_textures[n] = one_spectrum(ps, "value", 1.0f);
        } else if (is_float) {
          float t1 = tp_float(tp, "tex1", 1.0f), t2 = tp_float(tp, "tex2", 1.0f);
          if (cls == "scale") gs.float_textures[n] = t1 * t2;
          else { float amt = tp_float(tp, "amount", 0.5f); gs.float_textures[n] = t1 * (1.0f - amt) + t2 * amt; }
        } else {
          std::vector<float> t1 = tp_spectrum(tp, "tex1", spec.constant(1.0f)), t2 = tp_spectrum(tp, "tex2", spec.constant(1.0f));
          std::vector<float> r(t1.size());
          if (cls == "scale") { for (size_t k = 0; k < r.size(); k++) r[k] = t1[k] * t2[k]; }
          else { float amt = tp_float(tp, "amount", 0.5f); for (size_t k = 0; k < r.size(); k++) r[k] = t1[k] * (1.0f - amt) + t2[k] * amt; }
          gs.spec_textures[n] = r;
        }
      }
      else if (d == "LightSource") { std::string n = read_string(lx); ParamSet ps = parse_params(lx); light_source(n, ps); }
      else if (d == "AreaLightSource") { gs.area_light_name = read_string(lx); gs.area_light_params = parse_params(lx); }
      else if (d == "Shape") { std::string n = read_string(lx); ParamSet ps = parse_params(lx); shape(n, ps); }
      else if (d == "ObjectBegin") {
        std::string n = read_string(lx);
        gs_stack.push_back(gs); ctm_stack.push_back(ctm);
        hs->obj_prims.emplace_back();
        object_ids[n] = (int)hs->obj_prims.size() - 1;
        current_object = (int)hs->obj_prims.size() - 1;
      }
      else if (d == "ObjectEnd") {
        current_object = 0;
        if (gs_stack.empty()) throw ParseError("Unmatched ObjectEnd");
        gs = gs_stack.back(); gs_stack.pop_back();
        ctm = ctm_stack.back(); ctm_stack.pop_back();
      }
      else if (d == "ObjectInstance") {
        std::string n = read_string(lx);
        if (current_object != 0) continue;
        auto it = object_ids.find(n);
        if (it == object_ids.end()) { warn("ObjectInstance \"" + n + "\" unknown"); continue; }
        if (hs->obj_prims[(size_t)it->second].empty()) continue;
        pt_instance inst;
        inst.object = it->second;
        memcpy(inst.m, ctm.m.m, sizeof(inst.m));
        memcpy(inst.minv, ctm.minv.m, sizeof(inst.minv));
        hs->instances.push_back(inst);
        hs->obj_prims[0].push_back(~((int32_t)hs->instances.size() - 1));
      }
      else if (d == "Include") {
        std::string n = read_string(lx);
        std::string path = (n.size() && n[0] == '/') ? n : (dir.empty() ? n : dir + "/" + n);
        std::string text;
        if (!read_text_file(path, &text)) throw ParseError("Include: cannot open \"" + path + "\"");
        size_t slash = path.find_last_of('/');
        parse_text(text, slash == std::string::npos ? "" : path.substr(0, slash));
      }
      else if (d == "MakeNamedMedium") { read_string(lx); parse_params(lx); }
      else if (d == "MediumInterface") { read_string(lx); if (lx.peek().kind == TK_STRING) read_string(lx); }
      else throw ParseError("unknown directive '" + d + "'");
    }
  }

  // ---- filters (filters/*.rs) ------------------------------------------------------------------
  void make_filter(float radius[2], float table[256]) {
    const ParamSet& ps = filter_params;
    std::string n = filter_name;
    float dflt = 2.0f;
    if (n == "box") dflt = 0.5f; else if (n == "sinc") dflt = 4.0f;
    else if (n != "gaussian" && n != "mitchell" && n != "triangle") throw ParseError("Filter \"" + n + "\" unknown.");
    float rx = ps.one_float("xwidth", dflt), ry = ps.one_float("ywidth", dflt);
    radius[0] = rx; radius[1] = ry;
    float alpha = ps.one_float("alpha", 2.0f), B = ps.one_float("B", 1.0f / 3.0f), C = ps.one_float("C", 1.0f / 3.0f), tau = ps.one_float("tau", 3.0f);
    float exp_x = expf(-alpha * rx * rx), exp_y = expf(-alpha * ry * ry);
    float inv_rx = 1.0f / rx, inv_ry = 1.0f / ry;
    auto mitchell = [&](float x) {
      x = fabsf(2.0f * x);
      if (x > 1.0f) return ((-B - 6.0f * C) * x * x * x + (6.0f * B + 30.0f * C) * x * x + (-12.0f * B - 48.0f * C) * x + (8.0f * B + 24.0f * C)) * (1.0f / 6.0f);
      return ((12.0f - 9.0f * B - 6.0f * C) * x * x * x + (-18.0f + 12.0f * B + 6.0f * C) * x * x + (6.0f - 2.0f * B)) * (1.0f / 6.0f);
    };
    const float PI = 3.14159265358979323846f;
    auto sinc = [&](float x) { x = fabsf(x); return x < 1e-5f ? 1.0f : sinf(PI * x) / (PI * x); };
    auto wsinc = [&](float x, float r) { x = fabsf(x); return x > r ? 0.0f : sinc(x) * sinc(x / tau); };
    for (int y = 0; y < 16; y++)
      for (int x = 0; x < 16; x++) {  // film.rs:102-120 ("fixed" sampling x*r/15)
        float xx = (float)x * (rx / 15.0f), yy = (float)y * (ry / 15.0f);
        float v;
        if (n == "box") v = 1.0f;
        else if (n == "gaussian") v = fmaxf(0.0f, expf(-alpha * xx * xx) - exp_x) * fmaxf(0.0f, expf(-alpha * yy * yy) - exp_y);
        else if (n == "mitchell") v = mitchell(xx * inv_rx) * mitchell(yy * inv_ry);
        else if (n == "sinc") v = wsinc(xx, rx) * wsinc(yy, ry);
        else v = fmaxf(0.0f, rx - fabsf(xx)) * fmaxf(0.0f, ry - fabsf(yy));
        table[y * 16 + x] = v;
      }
  }

  // ---- finish: film, camera, sampler, integrator, accelerator ------------------------------------
  void finish() {
    pt_scene_desc& D = hs->desc;
    memset(&D, 0, sizeof(D));
    D.abi_version = PT_ABI_VERSION;
    D.spectrum_kind = (uint32_t)spec.kind();
    D.n_channels = (uint32_t)spec.n_channels();
    spec.xyz_weights(D.xyz_weights, &D.xyz_scale);

    // Film (film.rs:64-179, 538-571)
    pt_film& F = D.film;
    int xres = film_params.one_int("xresolution", 1280), yres = film_params.one_int("yresolution", 720);
    const pt_options& opt = pt_current_options();
    if (opt.quick_render && !opt.quick_render_full_resolution) {  // film.rs:548-554
      xres = std::max(1, xres / 4);
      yres = std::max(1, yres / 4);
    }
    float crop[4] = {0, 1, 0, 1};
    if (const Param* q = film_params.find(PC_FLOAT, "cropwindow")) {
      if (q->f.size() != 4) throw ParseError("\"cropwindow\" expects 4 values");
      for (int i = 0; i < 4; i++) crop[i] = q->f[(size_t)i];
    }
    F.full_resolution[0] = xres; F.full_resolution[1] = yres;
    // Bounds2f::from(((c0,c2),(c1,c3))): min = (c0,c2), max = (c1,c3)
    int x0 = std::max(0, (int)floorf((float)xres * crop[0])), y0 = std::max(0, (int)floorf((float)yres * crop[2]));
    int x1 = std::min((int)ceilf((float)xres * crop[1]), xres), y1 = std::min((int)ceilf((float)yres * crop[3]), yres);
    F.cropped_pixel_bounds[0] = x0; F.cropped_pixel_bounds[1] = y0; F.cropped_pixel_bounds[2] = x1; F.cropped_pixel_bounds[3] = y1;
    make_filter(F.filter_radius, F.filter_table);
    F.sample_bounds[0] = (int)floorf((float)x0 - F.filter_radius[0]);
    F.sample_bounds[1] = (int)floorf((float)y0 - F.filter_radius[1]);
    F.sample_bounds[2] = (int)ceilf((float)x1 + F.filter_radius[0]);
    F.sample_bounds[3] = (int)ceilf((float)y1 + F.filter_radius[1]);
    F.scale = film_params.one_float("scale", 1.0f);
    F.max_sample_luminance = film_params.one_float("maxsampleluminance", std::numeric_limits<float>::infinity());
    hs->filename = film_params.one_string("filename", "pbrt.exr");

    // Camera (perspective.rs:345-394, projective.rs:23-53)
    // orthographic.rs:193-253, environment.rs:90-151 read the same parameters (the environment camera ignores lens and window)
    if (camera_name != "perspective" && camera_name != "orthographic" && camera_name != "environment")
      throw ParseError("Camera \"" + camera_name + "\" is not supported on the GPU path (perspective, orthographic, environment)");
    pt_camera& Cm = D.camera;
    Cm.type = camera_name == "perspective" ? PT_CAMERA_PERSPECTIVE : (camera_name == "orthographic" ? PT_CAMERA_ORTHOGRAPHIC : PT_CAMERA_ENVIRONMENT);
    float so = camera_params.one_float("shutteropen", 0.0f), sc = camera_params.one_float("shutterclose", 1.0f);
    if (sc < so) std::swap(so, sc);
    float aspect = (float)xres / (float)yres;
    Cm.lens_radius = camera_params.one_float("lensradius", 0.0f);
    Cm.focal_distance = camera_params.one_float("focaldistance", 1e6f);
    Cm.shutter_open = so; Cm.shutter_close = sc;
    float frame = camera_params.one_float("frameaspectratio", aspect);
    float sw[4];  // min.x, max.x, min.y, max.y
    if (const Param* q = camera_params.find(PC_FLOAT, "screenwindow")) {
      if (q->f.size() < 4) throw ParseError("\"screenwindow\" should have four values");
      sw[0] = q->f[0]; sw[1] = q->f[1]; sw[2] = q->f[2]; sw[3] = q->f[3];
    } else if (frame > 1.0f) { sw[0] = -frame; sw[1] = frame; sw[2] = -1.0f; sw[3] = 1.0f; }
    else { sw[0] = -1.0f; sw[1] = 1.0f; sw[2] = -1.0f / frame; sw[3] = 1.0f / frame; }
    float fov = camera_params.one_float("fov", 90.0f);
    float halffov = camera_params.one_float("halffov", -1.0f);
    if (halffov > 0.0f) fov = 2.0f * halffov;
    Xf camera_to_screen = Cm.type == PT_CAMERA_PERSPECTIVE ? x_perspective(fov, 1e-2f, 1000.0f)
                                                           : x_scale(1.0f, 1.0f, 1.0f / (1.0f - 0.0f)) * x_translate(0.0f, 0.0f, -0.0f);  // Transform::orthographic(0, 1), transform.rs:85-87
    if (Cm.type == PT_CAMERA_ENVIRONMENT) Cm.lens_radius = 0.0f;
    Xf screen_to_raster = x_scale((float)xres, (float)yres, 1.0f) * x_scale(1.0f / (sw[1] - sw[0]), 1.0f / (sw[2] - sw[3]), 1.0f) *
                          x_translate(-sw[0], -sw[3], 0.0f);
    Xf raster_to_camera = xinverse(camera_to_screen) * xinverse(screen_to_raster);
    memcpy(Cm.raster_to_camera, raster_to_camera.m.m, sizeof(Cm.raster_to_camera));
    memcpy(Cm.camera_to_world, camera_to_world.m.m, sizeof(Cm.camera_to_world));

    // Sampler (samplers/*.rs create_*)
    pt_sampler& S = D.sampler;
    if (sampler_name == "halton") { S.type = PT_SAMPLER_HALTON; S.samples_per_pixel = sampler_params.one_int("pixelsamples", 16); }
    else if (sampler_name == "sobol") { S.type = PT_SAMPLER_SOBOL; S.samples_per_pixel = sampler_params.one_int("pixelsamples", 16); }
    else if (sampler_name == "random") { S.type = PT_SAMPLER_RANDOM; S.samples_per_pixel = sampler_params.one_int("pixelsamples", 4); }
    else if (sampler_name == "stratified") {  // create_stratified_sampler (stratified.rs:130-150)
      S.type = PT_SAMPLER_STRATIFIED;
      S.jitter = sampler_params.one_bool("jitter", true) ? 1 : 0;
      int xs = sampler_params.one_int("xsamples", 4), ys = sampler_params.one_int("ysamples", 4);
      S.n_sampled_dimensions = sampler_params.one_int("dimensions", 4);
      if (pt_current_options().quick_render) xs = ys = 1;
      // stratified_sample_2d loops x over 0..ny (sampling.rs:34-36): with xsamples != ysamples the reference leaves table entries
      // of the previous pixel in place (or indexes out of bounds); only the square case is reproduced
      if (xs != ys) throw ParseError("Sampler \"stratified\": xsamples != ysamples is not supported on the GPU path");
      if (xs < 1 || xs > 64 || S.n_sampled_dimensions < 0 || S.n_sampled_dimensions > 64) throw ParseError("Sampler \"stratified\": xsamples / dimensions out of range (1..64, 0..64)");
      S.strata = xs;
      S.samples_per_pixel = xs * ys;
    }
    else if (sampler_name == "02sequence" || sampler_name == "lowdiscrepancy") {  // create_zerotwosequence_sampler (zerotwosequence.rs:107-125)
      S.type = PT_SAMPLER_ZEROTWO;
      S.samples_per_pixel = sampler_params.one_int("pixelsamples", 16);
      S.n_sampled_dimensions = sampler_params.one_int("dimensions", 4);
      if (S.n_sampled_dimensions < 0 || S.n_sampled_dimensions > 64) throw ParseError("Sampler \"02sequence\": dimensions out of range (0..64)");
    }
    else throw ParseError("Sampler \"" + sampler_name + "\" is not supported on the GPU path (halton, sobol, random, stratified, 02sequence)");
    S.sample_at_pixel_center = sampler_params.one_bool("samplepixelcenter", false) ? 1 : 0;
    if (pixelsamples_override > 0 && S.type != PT_SAMPLER_STRATIFIED) S.samples_per_pixel = pixelsamples_override;  // pbrt.rs:234-238 (replace_params_int before the factory runs)
    if (pt_current_options().quick_render && (S.type < PT_SAMPLER_RANDOM || S.type == PT_SAMPLER_ZEROTWO)) S.samples_per_pixel = 1;  // halton.rs:288-290, sobol.rs:199-201

    // Integrator (integrators/path.rs:252-271)
    if (integrator_name != "path" && integrator_name != "aov" && integrator_name != "ambientocclusion" && integrator_name != "directlighting" && integrator_name != "whitted")
      throw ParseError("Integrator \"" + integrator_name + "\" is not supported on the GPU path (path, directlighting, whitted, ambientocclusion, aov)");
    pt_integrator& I = D.integrator;
    I.type = PT_INTEGRATOR_PATH;
    if (integrator_name == "ambientocclusion") {  // integrators/ao.rs:119-138
      I.type = PT_INTEGRATOR_AO;
      I.ao_cos_sample = integrator_params.one_bool("cossample", true) ? 1 : 0;
      I.ao_samples = integrator_params.one_int("nsamples", 64);
      if (pt_current_options().quick_render) I.ao_samples = 1;
      if (I.ao_samples < 1) throw ParseError("Integrator \"ambientocclusion\": nsamples must be positive");
      if (spec.n_channels() != 3) throw ParseError("Integrator \"ambientocclusion\" is only supported with RGB spectra");
    }
    if (integrator_name == "directlighting" || integrator_name == "whitted") {  // directlighting.rs:135-158, whitted.rs:109-130
      const bool direct = integrator_name == "directlighting";
      I.type = direct ? PT_INTEGRATOR_DIRECT : PT_INTEGRATOR_WHITTED;
      if (direct) I.direct_strategy = integrator_params.one_string("strategy", "all") == "one" ? PT_DIRECT_ONE : PT_DIRECT_ALL;
      if (!direct && integrator_params.find(PC_INT, "pixelbounds")) throw ParseError("Integrator \"whitted\": pixelbounds is not supported on the GPU path");
      // Sampler "random": the per-tile PCG32 streams are drawn from in the recursion's own order by the depth-first walk; the
      // "all" strategy's sample ARRAYS are drawn at every start_pixel (random.rs:46-60), before the pixel's first camera sample
      if (spec.n_channels() != 3) throw ParseError("Integrator \"" + integrator_name + "\" is only supported with RGB spectra");
    }
    if (integrator_name == "aov") {  // integrators/aov.rs:192-213, targets :36-58
      I.type = PT_INTEGRATOR_AOV;
      std::string t = integrator_params.one_string("target", "uv");
      I.aov_scale = integrator_params.one_float("scale", 1.0f);
      if (t == "distance") I.aov_target = PT_AOV_DISTANCE;
      else if (t == "depth") I.aov_target = PT_AOV_DEPTH;
      else if (t == "n" || t == "ng") I.aov_target = PT_AOV_N;
      else if (t == "ns" || t == "shading.n") I.aov_target = PT_AOV_NS;
      else if (t == "uv") I.aov_target = PT_AOV_UV;
      else if (t == "dpdus" || t == "shading.dpdu") I.aov_target = PT_AOV_DPDUS;
      else if (t == "rdxc" || t == "rdyc" || t == "drodx" || t == "drddx" || t == "dpdx" || t == "dpdy" || t == "dpdu" || t == "dpdv" || t == "dstdx" ||
               t == "dstdy" || t == "duvdx" || t == "dpdvs" || t == "shading.dpdv")
        throw ParseError("AOV target \"" + t + "\" is not supported on the GPU path (distance, depth, n, ns, uv, dpdus)");
      else throw ParseError("AOV target \"" + t + "\" unknown.");
      if (spec.n_channels() != 3) throw ParseError("Integrator \"aov\" is only supported with RGB spectra");
    }
    I.max_depth = integrator_params.one_int("maxdepth", 5);
    I.rr_threshold = integrator_params.one_float("rrthreshold", 1.0f);
    std::string ls = integrator_params.one_string("lightsamplestrategy", "spatial");
    I.light_strategy = ls == "uniform" ? PT_LIGHTS_UNIFORM : (ls == "power" ? PT_LIGHTS_POWER : PT_LIGHTS_SPATIAL);

    // Accelerator (create_bvh_accelerator.rs:9-28; accel/mod.rs:4-10)
    pt_accel& A = D.accel;
    if (accel_name != "bvh") throw ParseError("Accelerator \"" + accel_name + "\" is not supported on the GPU path (bvh only)");
    std::string layout = accel_params.one_string("layout", "qbvh");  // GPU-path extension: "lbvh" selects accel/lbvh.rs
    A.type = layout == "lbvh" ? PT_ACCEL_LBVH : PT_ACCEL_QBVH;
    std::string split = accel_params.one_string("splitmethod", "sah");
    A.split_method = split == "middle" ? PT_SPLIT_MIDDLE : (split == "equal" ? PT_SPLIT_EQUAL : (split == "hlbvh" ? PT_SPLIT_HLBVH : PT_SPLIT_SAH));
    A.max_prims_in_node = accel_params.one_int("maxnodeprims", 4);

    // emitters-only fix-up: they live after the sampled lights
    for (auto& li : hs->tri_light)
      if (li <= -2) li = (int32_t)hs->lights.size() + (-2 - li);
    std::vector<pt_light>& all = hs->lights;
    size_t n_sampled = all.size();
    all.insert(all.end(), hs->emitters_only.begin(), hs->emitters_only.end());

    // objects
    hs->obj_first.clear();
    hs->prims.clear();
    for (auto& v : hs->obj_prims) {
      hs->obj_first.push_back((uint32_t)hs->prims.size());
      hs->prims.insert(hs->prims.end(), v.begin(), v.end());
    }
    hs->obj_first.push_back((uint32_t)hs->prims.size());

    D.n_spectra = (uint32_t)(hs->spectra.size() / spec.n_channels());
    D.spectra = hs->spectra.data();
    D.n_tris = (uint32_t)hs->tri_flags.size();
    D.tri_p = hs->tri_p.data();
    if (!hs->tri_n.empty()) { hs->tri_n.resize((size_t)D.n_tris * 9, 0.0f); D.tri_n = hs->tri_n.data(); }
    if (!hs->tri_uv.empty()) { hs->tri_uv.resize((size_t)D.n_tris * 6, 0.0f); D.tri_uv = hs->tri_uv.data(); }
    if (!hs->tri_s.empty()) { hs->tri_s.resize((size_t)D.n_tris * 9, 0.0f); D.tri_s = hs->tri_s.data(); }
    D.tri_flags = hs->tri_flags.data();
    D.tri_material = hs->tri_material.data();
    D.tri_light = hs->tri_light.data();
    D.n_objects = (uint32_t)hs->obj_prims.size();
    D.obj_first = hs->obj_first.data();
    D.prims = hs->prims.data();
    D.n_instances = (uint32_t)hs->instances.size();
    D.instances = hs->instances.data();
    D.n_materials = (uint32_t)hs->materials.size();
    D.materials = hs->materials.data();
    D.n_lights = (uint32_t)n_sampled;
    D.n_emitters_only = (uint32_t)hs->emitters_only.size();
    D.n_spots = (uint32_t)hs->spots.size();
    D.spots = hs->spots.data();
    for (size_t k = 0; k < hs->envmaps.size(); k++) hs->envmaps[k].texels = hs->env_texels[k].data();
    D.n_envmaps = (uint32_t)hs->envmaps.size();
    D.envmaps = hs->envmaps.data();
    D.lights = all.data();
  }
};

// read_to_string (core/parser/read_file.rs:38-54): files whose extension is "gz" are gunzipped
bool read_text_file(const std::string& path, std::string* out) {
  out->clear();
  if (path.size() > 3 && path.substr(path.size() - 3) == ".gz") {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    char buf[1 << 16];
    int n;
    while ((n = gzread(g, buf, sizeof(buf))) > 0) out->append(buf, (size_t)n);
    gzclose(g);
    return n == 0;
  }
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}

// pbrt_parse_targz (core/parser/parser.rs:32-58): the archive is unpacked into a temporary directory, the first ".pbrt"
// file found in it is the scene, the directory is removed after the parse.  ustar / GNU archives with regular files,
// directories and GNU long names; entries that would leave the directory are refused.
struct TempDir {
  std::string path;
  ~TempDir() {
    if (path.empty()) return;
    nftw(path.c_str(), [](const char* p, const struct stat*, int, struct FTW*) { return remove(p); }, 16, FTW_DEPTH | FTW_PHYS);
  }
};
void mkdirs(const std::string& path) {
  for (size_t i = 1; i <= path.size(); i++)
    if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0700);
}
std::string untar_scene(const std::string& archive, TempDir* tmp) {
  std::string data;
  if (!read_text_file(archive, &data)) throw ParseError("cannot read \"" + archive + "\"");
  char templ[] = "/tmp/pbrt_r3_b200_XXXXXX";
  if (!mkdtemp(templ)) throw ParseError("cannot create a temporary directory");
  tmp->path = templ;
  std::string longname;
  for (size_t pos = 0; pos + 512 <= data.size();) {
    const char* h = data.data() + pos;
    if (h[0] == '\0') break;
    std::string name(h, strnlen(h, 100));
    if (memcmp(h + 257, "ustar", 5) == 0 && h[345]) name = std::string(h + 345, strnlen(h + 345, 155)) + "/" + name;
    const size_t size = (size_t)strtoull(std::string(h + 124, 12).c_str(), nullptr, 8);
    const char type = h[156];
    pos += 512;
    if (pos + size > data.size()) throw ParseError("truncated archive \"" + archive + "\"");
    if (type == 'L') { longname.assign(data.data() + pos, strnlen(data.data() + pos, size)); pos += (size + 511) / 512 * 512; continue; }
    if (type == 'x') {  // pax extended header: records "<length> key=value\n"; "path" names the next entry
      const std::string rec(data.data() + pos, size);
      for (size_t q = 0; q < rec.size();) {
        const size_t len = (size_t)strtoull(rec.c_str() + q, nullptr, 10);
        if (len == 0 || q + len > rec.size()) break;
        const std::string one = rec.substr(q, len);
        const size_t sp = one.find(' '), eq = one.find('=');
        if (sp != std::string::npos && eq != std::string::npos && one.substr(sp + 1, eq - sp - 1) == "path") longname = one.substr(eq + 1, one.size() - eq - 2);
        q += len;
      }
      pos += (size + 511) / 512 * 512;
      continue;
    }
    if (type == 'g') { pos += (size + 511) / 512 * 512; continue; }
    if (!longname.empty()) { name = longname; longname.clear(); }
    if (name.empty() || name[0] == '/' || name.find("..") != std::string::npos) throw ParseError("archive entry \"" + name + "\" leaves the archive");
    const std::string full = tmp->path + "/" + name;
    if (type == '5') mkdirs(full);
    else if (type == '0' || type == '\0') {
      size_t slash = full.find_last_of('/');
      mkdirs(full.substr(0, slash));
      std::ofstream o(full, std::ios::binary);
      o.write(data.data() + pos, (std::streamsize)size);
    }
    pos += (size + 511) / 512 * 512;
  }
  // the first .pbrt file, directories in name order (the reference takes read_dir's order)
  std::vector<std::string> found;
  std::vector<std::string> stack{tmp->path};
  while (!stack.empty() && found.empty()) {
    std::string d = stack.back();
    stack.pop_back();
    std::vector<std::string> names;
    if (DIR* dp = opendir(d.c_str())) {
      while (struct dirent* e = readdir(dp)) if (strcmp(e->d_name, ".") && strcmp(e->d_name, "..")) names.push_back(e->d_name);
      closedir(dp);
    }
    std::sort(names.begin(), names.end());
    std::vector<std::string> dirs;
    for (const std::string& n : names) {
      struct stat st;
      const std::string p = d + "/" + n;
      if (stat(p.c_str(), &st)) continue;
      if (S_ISDIR(st.st_mode)) dirs.push_back(p);
      else if (n.size() > 5 && n.substr(n.size() - 5) == ".pbrt") { found.push_back(p); break; }
    }
    for (auto it = dirs.rbegin(); it != dirs.rend(); ++it) stack.push_back(*it);
  }
  if (found.empty()) throw ParseError("no .pbrt file in \"" + archive + "\"");
  return found[0];
}

int parse_common(const std::string& text, const std::string& dir, int pixelsamples_override, int spectrum_kind, pt_host_scene** out) {
  *out = nullptr;
  std::unique_ptr<pt_host_scene> hs(new pt_host_scene());
  try {
    Builder b(hs.get(), spectrum_kind);
    b.pixelsamples_override = pixelsamples_override;
    b.parse_text(text, dir);
    if (!b.have_camera) b.camera_to_world = Xf::identity();
    b.finish();
  } catch (const std::exception& e) {
    pt_set_error(std::string("parse: ") + e.what());
    return 1;
  }
  *out = hs.release();
  return 0;
}

}  // namespace

extern "C" {

int pt_parse_string(const char* text, int pixelsamples_override, int spectrum_kind, pt_host_scene** out) {
  if (!text || !out) { pt_set_error("pt_parse_string: null argument"); return 1; }
  return parse_common(text, "", pixelsamples_override, spectrum_kind, out);
}

int pt_parse_file(const char* path, int pixelsamples_override, int spectrum_kind, pt_host_scene** out) {
  if (!path || !out) { pt_set_error("pt_parse_file: null argument"); return 1; }
  std::string p(path), text;
  TempDir tmp;  // a ".tar.gz" scene lives in a temporary directory for the duration of the parse (parser.rs:60-63)
  if (p.size() > 7 && p.substr(p.size() - 7) == ".tar.gz") {
    try { p = untar_scene(p, &tmp); } catch (const std::exception& e) { pt_set_error(std::string("pt_parse_file: ") + e.what()); return 1; }
  }
  if (!read_text_file(p, &text)) { pt_set_error(std::string("pt_parse_file: cannot open ") + p); return 1; }
  size_t slash = p.find_last_of('/');
  return parse_common(text, slash == std::string::npos ? "" : p.substr(0, slash), pixelsamples_override, spectrum_kind, out);
}

const pt_scene_desc* pt_host_scene_desc(const pt_host_scene* hs) { return hs ? &hs->desc : nullptr; }
const char* pt_host_scene_filename(const pt_host_scene* hs) { return hs ? hs->filename.c_str() : ""; }
uint32_t pt_host_scene_warning_count(const pt_host_scene* hs) { return hs ? (uint32_t)hs->warnings.size() : 0u; }
const char* pt_host_scene_warning(const pt_host_scene* hs, uint32_t i) { return (hs && i < hs->warnings.size()) ? hs->warnings[i].c_str() : nullptr; }
void pt_host_scene_destroy(pt_host_scene* hs) { delete hs; }

}  // extern "C"
