"""
``aggregated-classification`` module -- the direct consumer of nn-classification's NPZ; drop-in for
``genomad.aggregated_classification.main`` (reference genomad/modules/aggregated_classification.py:37-322): same
signature, same inputs required on disk, same files written, same skip/restart semantics and error behaviour.

The arithmetic (``branch_attention``, reference :10-34) is ~60 float64 operations per sequence on two [n, 3] score
matrices and one marker-frequency vector; it stays in NumPy float64 exactly as in the reference (a device kernel would
be launch latency and nothing else), and is held BIT-EXACT to the real reference function by golden vectors
(tests/golden/aggregate_golden.npz, made by tests/golden/make_golden.py importing the reference).

``aggregate_in_memory`` is the piece an ``end-to-end`` caller uses to skip the NPZ round trip between the two modules.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

from . import __version__, sequence, utils
from ._paths import AggregatedOutputs

_HEADER = "seq_name\tchromosome_score\tplasmid_score\tvirus_score\n"

# Trained constants of the attention-over-branches combiner (reference aggregated_classification.py:16-33).
_FREQ_GAIN = np.array([[0.3598502, 2.912244, -1.0668367, 1.3729712, -2.1972055, 0.9363847]])
_FREQ_OFFSET = np.array([[1.5372132, 2.6216774, -2.8225133, 3.0680428, 2.803005, -1.1982375]])
_MIX = np.array([[1.6666023, -1.1003100, -2.1425622],
                 [-2.2625937, 2.7540822, -1.5622343],
                 [1.9745151, 1.0952991, -2.7467837]])
_MIX_BIAS = np.array([0.14732242, -0.6838019, 0.5594167])


def softmax(x, temperature: float = 1.0, axis: int = 1):
    """reference utils.py:332-336 (max-subtracted, float64)."""
    x = np.asarray(x) / temperature
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / np.sum(e, axis=axis, keepdims=True)


def branch_attention(w, b1, b2, temperature: float = 2):
    """
    Combine the marker branch ``b1`` [n,3] and the neural-network branch ``b2`` [n,3] with gates that are an affine
    function of the total marker frequency ``w`` [n]:  alpha = w (x) gain + offset  (6 gates: 3 per branch);
    out = softmax((( b1*alpha[:, :3] + b2*alpha[:, 3:] ) / 2) @ M + c, T=2).   Operation order follows the reference
    so float64 results are bit-identical.
    """
    alpha = np.matmul(np.asarray(w).reshape(-1, 1), _FREQ_GAIN) + _FREQ_OFFSET
    g1 = b1 * alpha[:, 0:3]
    g2 = b2 * alpha[:, 3:6]
    return softmax(np.matmul((g1 + g2) / 2, _MIX) + _MIX_BIAS, temperature=temperature)


def total_marker_frequency(features: np.ndarray) -> np.ndarray:
    """Columns 15..17 of the marker feature matrix are the chromosome/plasmid/virus marker frequencies (reference :195)."""
    return features[:, 15:18].sum(1)


def aggregate_in_memory(features: np.ndarray, marker_predictions: np.ndarray, nn_predictions: np.ndarray) -> np.ndarray:
    """marker features [n,>=18] + the two branches' [n,3] scores -> aggregated float64 [n,3] (no files involved)."""
    return branch_attention(total_marker_frequency(features), marker_predictions, nn_predictions)


def _write_tsv(path: Path, names, preds) -> None:
    with open(path, "w") as fout:
        fout.write(_HEADER)
        for name, s in zip(names, preds):
            fout.write(f"{name}\t{s[0]:.4f}\t{s[1]:.4f}\t{s[2]:.4f}\n")


def main(input_path, output_path, restart, verbose):
    input_path, output_path = Path(input_path), Path(output_path)
    if not output_path.is_dir():
        output_path.mkdir()
    prefix = input_path.stem
    if sequence.is_compressed(input_path) != sequence.Compression.uncompressed:
        prefix = prefix.rsplit(".", 1)[0]
    outputs = AggregatedOutputs(prefix, output_path)
    console = utils.HybridConsole(output_file=outputs.aggregated_classification_log, verbose=verbose)
    parameter_dict = {}
    classify_proviruses = utils.check_provirus_execution(prefix, input_path, output_path)

    files = [outputs.aggregated_classification_execution_info, outputs.aggregated_classification_output,
             outputs.aggregated_classification_npz_output]
    descr = ["execution parameters", "sequence classification: tabular format", "sequence classification: binary format"]
    if classify_proviruses:
        files += [outputs.provirus_aggregated_classification_output, outputs.provirus_aggregated_classification_npz_output]
        descr += ["provirus classification: tabular format", "provirus classification: binary format"]
    utils.display_header(console, __version__, "aggregated-classification",
                         "This will aggregate the results of the marker-classification and nn-classification modules to "
                         "classify the input sequences into chromosome, plasmid, or virus.",
                         outputs.aggregated_classification_dir, files, descr)

    required = [outputs.marker_classification_execution_info, outputs.features_npz_output,
                outputs.marker_classification_npz_output, outputs.nn_classification_execution_info,
                outputs.nn_classification_npz_output]
    if classify_proviruses:
        required += [outputs.provirus_marker_classification_npz_output, outputs.provirus_nn_classification_npz_output]
    missing = [p.name for p in required if not p.exists()]
    if missing:
        console.error("The following files could not be found: " + ", ".join(missing) + ". Make sure to execute the "
                      "marker-classification and nn-classification modules.")
        sys.exit(1)

    input_md5 = utils.get_md5(input_path)
    if (input_md5 != utils.get_execution_info(outputs.marker_classification_execution_info)[0]
            or input_md5 != utils.get_execution_info(outputs.nn_classification_execution_info)[0]):
        console.error("Different input FASTA files were used as input for the marker-classification, nn-classification, "
                      "and aggregated-classification modules. Please execute all modules using the same input.")
        sys.exit(1)

    if not sequence.check_fasta(input_path):
        console.error(f"{input_path} is either empty or contains multiple entries with the same identifier. "
                      "Please check your input FASTA file and execute genomad aggregated-classification again.")
        sys.exit(1)
    console.log("Executing genomad aggregated-classification.")

    skip = False
    if outputs.aggregated_classification_execution_info.exists() and any(p.exists() for p in files) and not restart:
        if utils.compare_executions(input_path, parameter_dict, outputs.aggregated_classification_execution_info):
            skip = True
            console.log("Previous execution detected. Steps will be skipped unless their outputs are not found. "
                        "Use the --restart option to force the execution of all the steps again.")
        else:
            console.log("The input file or the parameters changed since the last execution. "
                        "Previous outputs will be overwritten.")
    if not outputs.aggregated_classification_dir.is_dir():
        console.log(f"Creating the {outputs.aggregated_classification_dir} directory.")
        outputs.aggregated_classification_dir.mkdir()
    utils.write_execution_info("aggregated_classification", input_path, parameter_dict,
                               outputs.aggregated_classification_execution_info)

    jobs = [("sequence", "Sequences", "contig_names", outputs.features_npz_output, "contig_features",
             outputs.marker_classification_npz_output, outputs.nn_classification_npz_output,
             outputs.aggregated_classification_npz_output, outputs.aggregated_classification_output)]
    if classify_proviruses:
        jobs.append(("provirus", "Proviruses", "provirus_names", outputs.provirus_features_npz_output,
                     "provirus_features", outputs.provirus_marker_classification_npz_output,
                     outputs.provirus_nn_classification_npz_output, outputs.provirus_aggregated_classification_npz_output,
                     outputs.provirus_aggregated_classification_output))
    # the reference loads every feature matrix before classifying anything (:192-204); a missing provirus feature file
    # therefore fails before any output is written -- keep that order
    freqs = [total_marker_frequency(np.load(j[3])[j[4]]) for j in jobs]
    console.log("The total marker frequencies of the input sequences were computed.")

    for (what, plural, names_key, _, _, marker_npz, nn_npz, out_npz, out_tsv), freq in zip(jobs, freqs):
        if skip and out_npz.exists():
            console.log(f"{out_npz.name} was found. Skipping {what} classification.")
            z = np.load(out_npz)
            names, preds = z[names_key], z["predictions"]
        else:
            m = np.load(marker_npz)
            names = m[names_key]
            preds = branch_attention(freq, m["predictions"], np.load(nn_npz)["predictions"])
            console.log(f"{plural} classified.")
            np.savez_compressed(out_npz, **{names_key: names, "predictions": preds})
            console.log(f"{what.capitalize()} classification in binary format written to {out_npz.name}.")
        _write_tsv(out_tsv, names, preds)
        console.log(f"{what.capitalize()} classification in tabular format written to {out_tsv.name}.")
    console.log("geNomad aggregated-classification finished!")
