"""Extract the reference's plugin surface (dataclass fields + method argument names) with `ast` — the modules import
nerfstudio at the top and cannot be imported here — into tests/golden/plugin_surface.json.
    python tests/golden/make_plugin_surface.py        (needs /root/reference; the JSON is committed)"""
import ast
import json
import os

REF = "/root/reference/signerf"
FILES = {"datasetgenerator/datasetgenerator.py": ["DatasetGeneratorConfig", "DatasetGenerator"],
         "diffuser/diffuser.py": ["DiffuserConfig", "Diffuser"],
         "renderer/renderer.py": ["RendererConfig", "Renderer"],
         "signerf_pipeline.py": ["SIGNeRFPipelineConfig", "SIGNeRFPipeline"]}
out = {}
for rel, classes in FILES.items():
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in classes:
            fields = [n.target.id for n in node.body if isinstance(n, ast.AnnAssign) and isinstance(n.target, ast.Name)]
            methods = {n.name: [a.arg for a in n.args.args] for n in node.body if isinstance(n, ast.FunctionDef)}
            out[node.name] = {"file": "signerf/" + rel, "fields": fields, "methods": methods}
# the two nerfstudio entry points (pyproject.toml:44-46)
import tomllib  # noqa: E402
out["entry_points"] = tomllib.load(open("/root/reference/pyproject.toml", "rb"))["project"]["entry-points"]["nerfstudio.method_configs"]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plugin_surface.json"), "w"), indent=1)
print({k: (len(v["fields"]), sorted(v["methods"])) for k, v in out.items() if k != "entry_points"})
