"""On-disk report of a Monte Carlo analysis in the reference's layout (monte_carlo.py:482-560): host I/O only.

Files: <dir>/monte_carlo_report.json, <dir>/monte_carlo_report.txt and <dir>/simulation_results/sim_<id>.json — the
per-sample dumps that the reference's find_max_apogee.py and analyze_outlier.py read.  A per-sample dump needs the full
time series, which the batch engine does not keep: each dumped sample is re-flown with the tape on
(`BatchRun.full_result`), so `max_samples` bounds the cost.
"""
from __future__ import annotations

import json
import os
from datetime import datetime

from .utils import object_to_serializable_dict, to_serializable

_METRICS = (("apogee_altitude", "Apogee Altitude", "m"), ("range", "Range", "m"), ("flight_time", "Flight Time", "s"))


def create_output_directory(root="outputs"):
    path = os.path.join(root, "monte_carlo_" + datetime.now().strftime("%Y%m%d_%H%M%S"))
    os.makedirs(path, exist_ok=True)
    return path


def save_report(analyzer, analysis, output_dir, max_samples=1000):
    os.makedirs(output_dir, exist_ok=True)
    counted = analysis["n_samples"] + analysis["n_failed"] + analysis["n_outliers"]
    report = {
        "timestamp": datetime.now().isoformat(),
        "simulation_summary": {
            "total_simulations": analysis["n_samples"], "failed_simulations": analysis["n_failed"],
            "outlier_simulations": analysis["n_outliers"],
            "success_rate": analysis["n_samples"] / counted * 100 if counted else float("nan"),
        },
        "uncertainty_parameters": analyzer.uncertainty_params,
        "parameter_ranges_observed": analysis.get("parameter_ranges_observed"),
    }
    for key, _, _ in _METRICS:
        report[key + "_stats"] = analysis[key]
    for name in ("rocket", "motor", "atmosphere", "wind_model"):
        report[name + "_parameters"] = object_to_serializable_dict(getattr(analyzer, name))
    if "performance" in analysis:
        report["performance"] = analysis["performance"]
    with open(os.path.join(output_dir, "monte_carlo_report.json"), "w") as fh:
        json.dump(to_serializable(report), fh, indent=2)

    sims = os.path.join(output_dir, "simulation_results")
    os.makedirs(sims, exist_ok=True)
    results = analysis.get("results", [])
    run = getattr(results, "_owner", None) or analyzer.last_run      # the run that produced THIS analysis
    for k in range(min(len(results), max_samples)):
        brief = results[k]
        sim_id = brief.get("simulation_id", k)
        full = run.full_result(int(sim_id) - run.first_id) if run is not None else brief
        full["simulation_id"] = sim_id
        with open(os.path.join(sims, f"sim_{sim_id}.json"), "w") as fh:
            json.dump(to_serializable(full), fh)

    lines = ["Monte Carlo Analysis Report", "=" * 50, "", f"Generated: {report['timestamp']}", "", "Simulation Summary:"]
    summ = report["simulation_summary"]
    lines += [f"  Valid simulations: {summ['total_simulations']}", f"  Failed simulations: {summ['failed_simulations']}",
              f"  Outlier simulations: {summ['outlier_simulations']}", f"  Success rate: {summ['success_rate']:.1f}%", ""]
    for key, title, unit in _METRICS:
        st = analysis[key]
        lines += [f"{title} Statistics:", f"  Mean: {st['mean']:.1f} {unit}", f"  Standard Deviation: {st['std']:.1f} {unit}",
                  f"  Min: {st['min']:.1f} {unit}", f"  Max: {st['max']:.1f} {unit}",
                  f"  95% Confidence Interval: [{st['percentiles'][0]:.1f}, {st['percentiles'][4]:.1f}] {unit}", ""]
    if "performance" in report:
        perf = report["performance"]
        lines += ["Performance Statistics:", f"  Total time: {perf['total_time']:.2f} s",
                  f"  Simulations per second: {perf['simulations_per_second']:.1f}", f"  Cores used: {perf['cores_used']}"]
    with open(os.path.join(output_dir, "monte_carlo_report.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return output_dir
