"""Assembles profiles/<tag>_summary.md from the files tools/capture.sh left in gpurun_out/ and copies the evidence into profiles/.
usage: python tools/make_summary.py <tag> "<one-line title>" """
import json
import os
import shutil
import subprocess
import sys

tag, title = sys.argv[1], sys.argv[2]
src, dst = "gpurun_out", "profiles"
parts = [f"# {title}\n",
         "`ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_run.py {msm,small}`, `... tools/ipa_time.py 14 1`, "
         "`... tools/ecfft_time.py 14` (cold-cache, serialised launches: compare shares, not absolutes; gen_points / at:: kernels are input "
         "generation).  Command list: `tools/capture.sh`.\n"]
for name, ttl in (("msm", "MSM 2^20 Pallas one-shot (device-resident inputs), 3 calls"), ("ntt", "NTT 2^20 Fp (best_fft), 4 calls"), ("small", "MSM 2^14+1 Vesta one-shot, 3 calls"),
                  ("ipa", "k=14 Vesta: Params setup (window tables), 1 commit, 14 IPA rounds"),
                  ("ecfft", "k=14 Vesta g -> g_lagrange (EC-iFFT + scale + batch_normalize), thread form then quad form, 4 calls each")):
    f = f"{src}/{tag}_launches_{name}.csv"
    if os.path.exists(f):
        shutil.copy(f, dst)
        parts.append(subprocess.run([sys.executable, "tools/summarize_launches.py", f, ttl], capture_output=True, text=True).stdout)
for extra in (f"{tag}_bench_n1.json", f"{tag}_ecfft_stage_ncu_full_raw.csv", f"{tag}_pytest_gpu.txt", f"{tag}_smoke.txt", f"{tag}_replay_time.txt"):
    if os.path.exists(f"{src}/{extra}") and os.path.getsize(f"{src}/{extra}"):
        shutil.copy(f"{src}/{extra}", dst)
b = f"{src}/{tag}_bench_n1.json"
if os.path.exists(b) and os.path.getsize(b):
    d = json.loads(open(b).read().strip().splitlines()[-1])
    r, e, x = d["roofline"], d["e2e"], d["extra"]
    parts.append("### bench.py line (1 GPU)\n")
    parts.append(f"* MSM 2^20 device-resident: {d['value'] / 1e6:.1f} M pairs/s ({d['ms_per_step']:.3f} ms); through the host API: "
                 f"{e['value'] / 1e6:.1f} M pairs/s ({e['ms_per_step']:.3f} ms); CPU restatement {d['cpu_baseline']['value'] / 1e6:.2f} M pairs/s "
                 f"on {d['cpu_baseline']['cores']} threads.")
    parts.append(f"* {r['kernel']}: {r['kernel_ms']:.3f} ms per launch, {r['achieved']:.1f} GB/s algorithmic = {100 * r['frac']:.2f} % of the "
                 f"{r['peak']} GB/s HBM peak; {r['compute']['achieved']:.1f} of {r['compute']['peak']:.1f} G modmul/s ({100 * r['compute']['frac']:.0f} %).")
    n = x["ntt"]
    parts.append(f"* NTT 2^20: {n['value'] / 1e9:.2f} G elems/s ({n['ms_per_step']:.3f} ms), host API {n['e2e']['value'] / 1e9:.2f} G elems/s.")
    c = x["create_proof_k14_replay"]
    if "value" in c:
        parts.append(f"* create_proof k=14 proof-shaped replay (Blake2b transcript in both arms, transcript_identical = {c.get('transcript_identical')}): "
                     f"{c['value']:.2f} ms vs {c['cpu_baseline']['value']:.0f} ms on the C restatement ({c['cpu_baseline']['cores']} threads) = "
                     f"{c['cpu_baseline']['value'] / c['value']:.1f}x; by kind (synchronised after every call) "
                     f"{json.dumps({k: round(v, 3) for k, v in c.get('gpu_ms_by_kind_synced', {}).items()})}; Params setup {c['params_setup_ms']:.0f} ms.")
    e2 = d["e2e"]
    parts.append(f"* e2e by caller memory: pageable (staged, headline) {e2['ms_per_step']:.3f} ms, pinned {e2['pinned']['ms_per_step']:.3f} ms, "
                 f"pageable with the staging ring off {e2['pageable_plain']['ms_per_step']:.3f} ms.")
    if "msm_2p24_strong" in x and x["msm_2p24_strong"] and "value" in x["msm_2p24_strong"]:
        m24 = x["msm_2p24_strong"]
        parts.append(f"* MSM 2^24 (BASELINE configs[4], N = {m24['n_gpus']}): {m24['value'] / 1e6:.1f} M pairs/s ({m24['ms_per_step']:.2f} ms), parity vs oracle {m24.get('parity_vs_oracle')}.")
    if "lookup_permute_k14" in x and "gpu_ms" in x["lookup_permute_k14"]:
        lp = x["lookup_permute_k14"]
        parts.append(f"* lookup permuted columns at k=14: {lp['gpu_ms']:.3f} ms vs {lp['cpu_baseline']['ms']:.1f} ms (C restatement, 1 core), same result: {lp['same_result']}.")
    if "params_lagrange_k14" in x:
        p = x["params_lagrange_k14"]
        parts.append(f"* g -> g_lagrange at k=14: {p['gpu_ms']:.2f} ms vs CPU restatement {p['cpu_baseline']['ms']:.0f} ms at k={p['cpu_baseline']['k']} "
                     f"({p['cpu_baseline']['cores']} threads).")
    if "quotient_pipeline_k14" in x and "gpu_ms" in x["quotient_pipeline_k14"]:
        q = x["quotient_pipeline_k14"]
        parts.append(f"* resident quotient pipeline at k=14 (4 columns -> extended, {q['ast_instructions']}-instruction Ast, / vanishing, -> coefficients): "
                     f"{q['gpu_ms']:.3f} ms vs {q['cpu_baseline']['ms']:.0f} ms on the C restatement ({q['cpu_baseline']['cores']} threads), same result: {q['same_result']}.")
    if "poly_reductions_k14" in x and "eval_x16" in x["poly_reductions_k14"]:
        r_ = x["poly_reductions_k14"]
        parts.append(f"* k=14 reductions on resident polynomials: 16 eval_polynomial {r_['eval_x16']['gpu_ms']:.3f} ms vs {r_['eval_x16']['cpu_baseline']['ms']:.1f} ms, "
                     f"4 kate_division {r_['kate_division_x4']['gpu_ms']:.3f} ms vs {r_['kate_division_x4']['cpu_baseline']['ms']:.1f} ms (1 core).")
    if "verify" in c:
        v = c["verify"]
        parts.append(f"* verification of that proof (multiopen MSM, the opening, compute_s on the device, ONE multiexp over the resident generators): "
                     f"{v['value']:.2f} ms wall through the Python mirror vs {v['cpu_baseline']['value']:.1f} ms of hot calls on the C restatement; accepted = "
                     f"{v['accepted']}, tampered rejected = {v['tampered_rejected']}.")
    c16 = x.get("create_proof_k16_replay", {})
    if "value" in c16:
        parts.append(f"* k=16 replay: {c16['value']:.2f} ms vs {c16['cpu_baseline']['value']:.0f} ms = {c16['cpu_baseline']['value'] / c16['value']:.1f}x (transcript_identical = "
                     f"{c16.get('transcript_identical')}); its verification {c16.get('verify', {}).get('value', float('nan')):.2f} ms vs "
                     f"{c16.get('verify', {}).get('cpu_baseline', {}).get('value', float('nan')):.1f} ms.")
    gp = x.get("golden_proofs_verify_k11", {})
    if "accepted" in gp:
        parts.append(f"* the reference's {gp['proofs']} stored k=11 proofs through the engine: accepted = {gp['accepted']}, tampered rejected = {gp['tampered_rejected']}; "
                     f"{gp['gpu_ms_per_proof_wall']:.1f} ms wall per proof (the Python restatement of plonk::verify_proof dominates), the path's tail "
                     f"(use_challenges + eval) {gp['gpu_ms_per_proof_path_tail']:.2f} ms vs {gp['cpu_baseline']['ms_per_proof_hot']:.2f} ms of hot calls on the C restatement; "
                     f"BatchVerifier shape: {gp['batch']['gpu_ms_accumulate_and_eval']:.1f} ms to accumulate all {gp['proofs']} MSMs and evaluate once (accepted = {gp['batch']['accepted']}).")
    parts.append(f"* clocks {d['clocks']}\n")
open(f"{dst}/{tag}_summary.md", "w").write("\n".join(parts))
print(open(f"{dst}/{tag}_summary.md").read()[:3000])
