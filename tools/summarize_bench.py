"""Markdown rows from bench.py JSON lines:  python tools/summarize_bench.py profiles/r2_final_*.json"""
import json
import sys


def row(path):
    d = json.load(open(path))
    if d.get("impl") == "reference":
        cb = d["cpu_baseline"]
        return (f"| {d['metric']} (reference arm, {cb['kind']}) | {d['config']['workload']} | {d['value']:.2f} | - | - | - | "
                f"{cb['cores']} host cores | - |")
    r = d["roofline"]
    cpu = d.get("cpu_baseline") or {}
    par = d.get("parity_vs_cpu") or {}
    ok = par.get("ok")
    return (f"| {d['metric']} | {d['config']['workload'][:70]} | {d['value']:.1f} | {d['e2e']['value']:.1f} | "
            f"{d.get('aligned_words_per_s', 0):.0f} | {r['kernel'].split(' ')[0]} {r['achieved']:.0f} {r['unit']} = {r['frac']:.2f} ({r['bound']}) | "
            f"{cpu.get('value', float('nan')):.2f} ({cpu.get('cores', '?')} cores, {cpu.get('kind', '?')}) | {ok} |")


if __name__ == "__main__":
    print("| metric | workload | value (device-timed) | e2e | words/s | roofline of the dominant kernel | CPU baseline audio-s/s | parity_vs_cpu |")
    print("|---|---|---|---|---|---|---|---|")
    for p in sys.argv[1:]:
        try:
            print(row(p))
        except Exception as e:
            print(f"| {p} | unreadable: {e} | | | | | | |")
