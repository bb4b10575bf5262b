####################################################
# exp_type_2.smk -- B200 drop-in for workflow/rules/exp_type_2.smk of vshiv18/khoice (include it instead of
# rules/exp_type_2.smk at workflow/Snakefile:53).
#
# The UNMODIFIED reference file already runs on the B200 engine rule by rule when khoice_b200/bin is first on PATH
# (kmc, kmc_tools transform / complex / simple ... intersect -ocsum / kmers_subtract are served by khoice_b200/cli.py).
# This file is the fused alternative: the two table rules keep their names and outputs, and ONE job produces every
# histogram they read (khoice_b200/pipeline2.py: per (k, dataset) one sort of rest-of-set + pivot windows with the
# genome id as payload, per k one sort of all unions and pivot sets).  Intermediate databases are header-only stubs.
# Run with `--cores 1` per GPU.
####################################################
import os

if "K_VALUES" in config:
    k_values = [str(x) for x in str(config["K_VALUES"]).split(",")]

if exp_type == 2:
    from khoice_b200 import pipeline2 as khb_p2, tables as khb_tables
    if not os.path.isdir("input_type_2"):
        khb_p2.prepare_inputs(".", database_root, curr_trial, num_datasets)           # reference :31-48
    khb_p2.write_complex_ops(".", k_values, num_datasets)                              # reference :27-29, 50-113

rule pivot_histograms_exp_type2:
    """rules build_kmc_database_on_{genome,pivot}_exp_type_2 .. across_group_histogram_exp_type2 of the reference, fused"""
    input:
        lambda w: [khb_p2.p_pivot(n) for n in range(1, num_datasets + 1)]
    output:
        khb_p2.histogram_files("within", k_values, num_datasets) if exp_type == 2 else [],
        khb_p2.histogram_files("across", k_values, num_datasets) if exp_type == 2 else []
    run:
        rep = khb_p2.run_fused(".", num_datasets, k_values)
        # run_fused also writes the two CSVs; the rules below own them, so they are rebuilt there
        for f in (khb_p2.P_WITHIN_CSV, khb_p2.P_ACROSS_CSV):
            if os.path.exists(f):
                os.remove(f)

rule within_group_analysis_exp_type2:
    input:
        khb_p2.histogram_files("within", k_values, num_datasets) if exp_type == 2 else []
    output:
        "within_dataset_analysis_type_2/within_dataset_analysis.csv"
    run:
        khb_tables.within_group_analysis_exp_type2(list(input), output[0], num_datasets)

rule across_group_analysis_exp_type2:
    input:
        khb_p2.histogram_files("across", k_values, num_datasets) if exp_type == 2 else []
    output:
        "across_dataset_analysis_type_2/across_dataset_analysis.csv"
    run:
        khb_tables.across_group_analysis_exp_type2(list(input), output[0], num_datasets)

rule generate_exp2_output:
    input:
        "within_dataset_analysis_type_2/within_dataset_analysis.csv",
        "across_dataset_analysis_type_2/across_dataset_analysis.csv"
