"""Run the kernel/UNet parity checks of tests/kernel_checks.py and log one JSON line per check to
gpurun_out/gpu_check.jsonl.  Checks run in a worker process; if a kernel traps (which poisons the CUDA
context) the worker is restarted on the remaining checks, so one bad kernel only loses its own check.
Usage (GPU box):  python tools/gpu_check.py [name-substring ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def worker(names) -> None:
    import torch

    import kernel_checks as kc
    table = {**kc.ALL_CHECKS, **kc.UNET_CHECKS}
    for n in names:
        print("START " + n, flush=True)
        t0 = time.time()
        try:
            r = table[n]()
        except Exception as e:  # noqa: BLE001
            r = dict(ok=False, error=f"{type(e).__name__}: {e}"[:800])
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001 - context is gone; let the parent restart us
                r["name"] = n
                print("RESULT " + json.dumps(r), flush=True)
                sys.exit(17)
        r["name"] = n
        r["secs"] = round(time.time() - t0, 2)
        print("RESULT " + json.dumps(r), flush=True)


def main() -> None:
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(sys.argv[2:])
        return
    import kernel_checks as kc
    names = list(kc.ALL_CHECKS) + list(kc.UNET_CHECKS)
    if len(sys.argv) > 1:
        names = [n for n in names if any(s in n for s in sys.argv[1:])]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.jsonl"), "a")
    results = {}
    todo = list(names)
    while todo:
        p = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker"] + todo, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
        try:
            out, err = p.communicate(timeout=180 + 60 * len(todo))
        except subprocess.TimeoutExpired:
            p.kill()
            out, err = p.communicate()
        started = None
        for line in out.splitlines():
            if line.startswith("START "):
                started = line[6:]
            elif line.startswith("RESULT "):
                rec = json.loads(line[7:])
                results[rec["name"]] = rec
                started = None
        if started is not None and started not in results:      # died inside this check
            results[started] = dict(name=started, ok=False, error="worker died", rc=p.returncode,
                                    stderr=err[-1200:], stdout="\n".join(out.splitlines()[-8:])[-800:])
        done = set(results)
        remaining = [n for n in todo if n not in done]
        if len(remaining) == len(todo):                           # no progress: give up on the first
            results[todo[0]] = dict(name=todo[0], ok=False, error="worker made no progress", stderr=err[-1200:])
            remaining = todo[1:]
        todo = remaining
    n_ok = 0
    for n in names:
        rec = results[n]
        n_ok += bool(rec.get("ok"))
        log.write(json.dumps(rec) + "\n")
        print(("PASS " if rec.get("ok") else "FAIL ") + json.dumps(rec)[:700], flush=True)
    log.close()
    print(f"{n_ok}/{len(names)} checks passed")


if __name__ == "__main__":
    main()
