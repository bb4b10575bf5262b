####################################################
# exp_type_4.smk -- B200 drop-in for workflow/rules/exp_type_4.smk of vshiv18/khoice (include it instead of
# rules/exp_type_4.smk at workflow/Snakefile:55).
#
# The UNMODIFIED reference file already runs on the B200 engine rule by rule when khoice_b200/bin is first on PATH
# (kmc, kmc_tools transform ... set_counts / histogram / dump -s, complex, simple ... intersect -ocsum), and its
# src/merge_lists.py works unchanged on the text dumps the shims write.  This file is the fused alternative
# (khoice_b200/pipeline4.py): per k ONE job computes every pivot k-mer's group-membership bitmask on the GPU and
# accumulates the confusion matrix in the reference's order of additions; no text dumps, no G x G intersections.
# Feature level only (the reference's rule does not pass -r).  Run with `--cores 1` per GPU.
####################################################
import os

if "K_VALUES" in config:
    k_values = [str(x) for x in str(config["K_VALUES"]).split(",")]

if exp_type == 4:
    from khoice_b200 import pipeline4 as khb_p4
    if not os.path.isdir("input_type4"):
        khb_p4.prepare_inputs(".", database_root, curr_trial, num_datasets, out_pivot=bool(out_pivot_exp4))   # reference :31-52
    khb_p4.write_parse_time_files(".", k_values, num_datasets)                                               # reference :27-29, 54-103

rule run_merge_list_exp_type_4:
    """rules build_kmc_database_on_*_exp_type_4 .. run_merge_list_exp_type_4 of the reference for one k, fused"""
    input:
        lambda w: [khb_p4.p_pivot(n) for n in range(1, num_datasets + 1)]
    output:
        "accuracies_type_4/values/k_{k}_accuracy_values.csv",
        "accuracies_type_4/confusion_matrix/k_{k}_confusion_matrix.txt"
    run:
        khb_p4.run_fused(".", num_datasets, [wildcards.k])

rule concatenate_accuracies_exp_type4:
    input:
        expand("accuracies_type_4/values/k_{k}_accuracy_values.csv", k=k_values)
    output:
        "accuracies_type_4/accuracy_values.csv"
    shell:
        "cat accuracies_type_4/values/*.csv > accuracies_type_4/accuracy_values.csv"

rule generate_exp4_output:
    input:
        "accuracies_type_4/accuracy_values.csv"
