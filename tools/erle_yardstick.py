"""ERLE of the frozen STFT-domain recurrence (numpy oracle, float64) next to the classical
overlap-save constrained PBFDAF (oracle/pbfdaf_yardstick.py) on the SURVEY 8d sets.  CPU only; analysis, not product.

    python tools/erle_yardstick.py [--utterances 8] [--seconds 10]

single talk : ERLE = sum mic^2 / sum err^2 over the last 60 % of the utterance
double talk : echo-only ERLE = sum echo^2 / sum (echo - yhat)^2 over the double-talk span [0.4 L, 0.7 L] and after it
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acoustic_echo_cancellation_b200 import synth  # noqa: E402
from oracle import aec_oracle as O  # noqa: E402
from oracle.pbfdaf_yardstick import pbfdaf  # noqa: E402


def db(num, den):
    return 10.0 * np.log10(max(float(np.sum(num ** 2)), 1e-20) / max(float(np.sum(den ** 2)), 1e-20))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=10.0)
    args = ap.parse_args()
    L = int(args.seconds * 16000)
    rows = []
    for P, algo, name in [(4, O.ALGO_NLMS, "P=4 NLMS"), (4, O.ALGO_KALMAN, "P=4 Kalman"), (8, O.ALGO_NLMS, "P=8 NLMS"),
                          (8, O.ALGO_KALMAN, "P=8 Kalman"), (16, O.ALGO_KALMAN, "P=16 Kalman"), (16, O.ALGO_NLMS, "P=16 NLMS")]:
        for dt in (False, True):
            acc = {"stft": [], "stft_dt": [], "pb": [], "pb_dt": [], "pb_unc": [], "algo2": [], "algo2_dt": [], "algo3": [],
                   "algo3_dt": []}
            for u in range(args.utterances):
                d = synth.make_utterance(u, L, rir_len=P * 256, double_talk=dt)
                far, mic, echo = (d[k].astype(np.float64) for k in ("far", "mic", "echo"))
                r = O.stage1(far[None], mic[None], O.AecConfig(partitions=P, algo=algo))
                e, yh = r["err"][0], r["echo"][0]
                pe, py = pbfdaf(far, mic, P)
                pu, _ = pbfdaf(far, mic, P, constrained=False)
                n = min(len(e), len(pe))
                # the product's overlap-save filters (algos 2 / 3: alternated constraint on the partition as it entered the
                # block, NLMS step on a smoothed power / Kalman step) -- stated once per filter length, not per STFT step rule
                for key, a2 in (("algo2", O.ALGO_PBFDAF), ("algo3", O.ALGO_PBFKF)):
                    if algo != O.ALGO_NLMS:
                        continue
                    r2 = O.stage1(far[None], mic[None], O.AecConfig(partitions=P, algo=a2))
                    e2, y2 = r2["err"][0], r2["echo"][0]
                    if not dt:
                        lo = int(0.4 * n)
                        acc[key].append(db(mic[lo:n], e2[lo:n]))
                    else:
                        a, b = int(0.4 * L), int(0.7 * L)
                        acc[key + "_dt"].append(db(echo[a:b], (echo[:n] - y2[:n])[a:b]))
                        acc[key].append(db(echo[b:n], (echo[:n] - y2[:n])[b:n]))
                if not dt:
                    lo = int(0.4 * n)
                    acc["stft"].append(db(mic[lo:n], e[lo:n]))
                    acc["pb"].append(db(mic[lo:n], pe[lo:n]))
                    acc["pb_unc"].append(db(mic[lo:n], pu[lo:n]))
                else:
                    a, b = int(0.4 * L), int(0.7 * L)
                    acc["stft_dt"].append(db(echo[a:b], (echo[:n] - yh[:n])[a:b]))
                    acc["pb_dt"].append(db(echo[a:b], (echo[:n] - py[:n])[a:b]))
                    acc["stft"].append(db(echo[b:n], (echo[:n] - yh[:n])[b:n]))
                    acc["pb"].append(db(echo[b:n], (echo[:n] - py[:n])[b:n]))
            row = {"filter": name, "set": "double talk (SER 0 dB)" if dt else "single talk",
                   "utterances": args.utterances, "seconds": args.seconds}
            for k, v in acc.items():
                if v:
                    row[k + "_erle_db_mean"] = round(float(np.mean(v)), 2)
            rows.append(row)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
