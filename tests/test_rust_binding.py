"""The Rust side of the boundary (rust/build.rs, rust/src/batch.rs) cannot be compiled in this image (no cargo / rustc),
so its `extern "C"` block is checked against include/bbs_b200.h prototype by prototype: same names, same arity, same
types in the same order, and no header symbol left unbound.  Also: build.rs names the real source files, and the status
mapping does not fold `malformed` into a reference error variant (round-1 defect)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_TO_RUST = {
    "int": "i32", "size_t": "usize", "uint32_t": "u32", "uint64_t": "u64", "float": "f32", "double": "f64",
    "const uint8_t*": "*const u8", "uint8_t*": "*mut u8", "const uint32_t*": "*const u32", "uint32_t*": "*mut u32",
    "const uint64_t*": "*const u64", "uint64_t*": "*mut u64", "float*": "*mut f32", "double*": "*mut f64",
    "void*": "*mut c_void", "const char*": "*const c_char", "void": "()",
    "bbs_ctx*": "*mut BbsCtx", "bbs_ctx**": "*mut *mut BbsCtx",
    "bbs_issuer_set*": "*mut BbsIssuerSet", "bbs_issuer_set**": "*mut *mut BbsIssuerSet",
}


def c_prototypes():
    text = open(os.path.join(ROOT, "include", "bbs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(bbs_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if params and params != "void":
            for p in params.split(","):
                p = re.sub(r"(\w+)\s*\[[^\]]*\]", r"* \1", p.strip())             # array parameter = pointer
                p = re.sub(r"\s*\*\s*", "* ", p)
                toks = p.split()
                if len(toks) > 1 and not toks[-1].endswith("*"):
                    toks = toks[:-1]                                           # drop the parameter name
                t = " ".join(toks).replace(" *", "*").replace("* *", "**").replace("* ", "*").strip()
                types.append(t)
        out[name] = (ret.replace(" *", "*"), types)
    return out


def rust_prototypes():
    text = open(os.path.join(ROOT, "rust", "src", "batch.rs")).read()
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    block = re.sub(r"//.*$", "", block, flags=re.M)
    out = {}
    for m in re.finditer(r"fn\s+(bbs_[a-z0-9_]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block, flags=re.S):
        name, params, ret = m.group(1), m.group(2), (m.group(3) or "()").strip()
        types = [" ".join(p.split(":", 1)[1].split()) for p in params.split(",") if p.strip()]
        out[name] = (ret, types)
    return out


def test_extern_block_matches_the_header():
    c, r = c_prototypes(), rust_prototypes()
    assert len(c) >= 40
    assert sorted(c) == sorted(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, (cret, ctypes_) in c.items():
        rret, rtypes = r[name]
        assert C_TO_RUST[cret] == rret, (name, cret, rret)
        assert [C_TO_RUST[t] for t in ctypes_] == rtypes, (name, ctypes_, rtypes)


def test_python_binding_matches_the_header_arity():
    from bbs_sign_b200 import _native
    c = c_prototypes()
    assert sorted(c) == sorted(_native.SYMBOLS)
    for name, (_, types) in c.items():
        assert len(_native.SYMBOLS[name][1]) == len(types), name


def test_build_rs_names_real_sources():
    text = open(os.path.join(ROOT, "rust", "build.rs")).read()
    groups = re.search(r"const GROUPS: \[&str; (\d+)\] = \[([^\]]*)\]", text)
    names = re.findall(r'"([a-z0-9]+)"', groups.group(2))
    assert int(groups.group(1)) == len(names)
    csrc = os.path.join(ROOT, "bbs_sign_b200", "csrc")
    have = sorted(f[3:-3] for f in os.listdir(csrc) if f.startswith("tu_") and f.endswith(".cu"))
    assert sorted(names) == have                   # the same translation units as the Makefile
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert sorted(re.search(r"GROUPS\s*:=\s*(.*)", mk).group(1).split()) == have
    assert "arch=compute_100a,code=sm_100a" in text and "capi.cu" in text


def test_status_mapping_keeps_malformed_distinct():
    text = open(os.path.join(ROOT, "rust", "src", "batch.rs")).read()
    sig = re.search(r"fn signature_status\(s: u8\).*?\n\}", text, flags=re.S).group(0)
    assert "ST_ERR_MSG_GEN_LEN => Err(ItemError::Reference(SignatureError::InvalidMessageAndGeneratorsLength))" in sig
    assert "_ => Err(ItemError::Malformed)" in sig
    hdr = open(os.path.join(ROOT, "include", "bbs_b200.h")).read()
    for name, val in re.findall(r"const (ST_[A-Z_]+): u8 = (\d+);", text):
        assert re.search(rf"#define BBS_{name} {val}\b", hdr), name
    assert "impl CurveId for Bls12381Const" in text and "impl CurveId for Bn254Const" in text
