"""Writes tests/golden/reference_signatures.json: names, parameter order and defaults of the reference
functions this repo mirrors, parsed (not imported - its dependencies are absent) from /root/reference.
Run in the build container: `python tests/golden/make_signatures.py`."""
import ast
import json
import os

REF = "/root/reference/keypoint_moseq"
WANT = {
    "fitting.py": ["fit_model", "apply_model", "estimate_syllable_marginals", "update_hypparams",
                   "expected_marginal_likelihoods", "init_model"],
    "io.py": ["save_hdf5", "load_hdf5", "load_checkpoint", "reindex_syllables_in_checkpoint", "extract_results",
              "load_results"],
    "util.py": ["format_data"],
}
out = {}
for fname, names in WANT.items():
    tree = ast.parse(open(os.path.join(REF, fname)).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            a = node.args
            pos = [x.arg for x in a.args]
            defaults = [None] * (len(pos) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
            params = [{"name": n, "default": d, "required": d is None and i < len(pos) - len(a.defaults)}
                      for i, (n, d) in enumerate(zip(pos, defaults))]
            for x, d in zip(a.kwonlyargs, a.kw_defaults):
                params.append({"name": x.arg, "default": None if d is None else ast.unparse(d), "required": d is None})
            out[f"{fname[:-3]}.{node.name}"] = {"params": params, "varargs": a.vararg.arg if a.vararg else None,
                                                "varkw": a.kwarg.arg if a.kwarg else None, "line": node.lineno}
# default hyper-parameters: the dict literals inside generate_config (io.py:46-169)
tree = ast.parse(open(os.path.join(REF, "io.py")).read())
gen = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "generate_config")
defaults = {}
for node in ast.walk(gen):
    if isinstance(node, ast.Dict):
        for k, v in zip(node.keys, node.values):
            if isinstance(k, ast.Constant) and k.value in ("error_estimator", "obs_hypparams", "ar_hypparams",
                                                           "trans_hypparams", "cen_hypparams"):
                defaults[k.value] = ast.literal_eval(v)
            if isinstance(k, ast.Constant) and k.value in ("conf_pseudocount", "whiten", "fix_heading",
                                                           "added_noise_level", "PCA_fitting_num_frames",
                                                           "conf_threshold") and isinstance(v, ast.Constant):
                defaults[k.value] = v.value
out["__generate_config_defaults__"] = defaults
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json"), "w"), indent=1)
print({k: len(v["params"]) for k, v in out.items() if "params" in v})
print(defaults)
