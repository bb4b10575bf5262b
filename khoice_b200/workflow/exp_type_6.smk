####################################################
# exp_type_6.smk -- B200 drop-in for workflow/rules/exp_type_6.smk of vshiv18/khoice (include it instead of
# rules/exp_type_6.smk in workflow/Snakefile).
#
# The UNMODIFIED reference file already runs on the B200 engine rule by rule when khoice_b200/bin is first on PATH
# (kmc, kmc_tools transform ... set_counts / histogram / dump -s, complex, simple ... intersect -ocsum), and its
# src/merge_lists.py works unchanged on the text dumps the shims write.  This file is the fused alternative
# (khoice_b200/pipeline6.py): per k ONE job computes, for both read types, every read k-mer's group-membership bit mask and
# every read's votes on the GPU (csrc/reads.cu) and draws the confusion matrix with the reference's tie-breaking; no text
# dumps, no G x G intersections.  Read level, like the reference's rule (it passes -r).  Run with `--cores 1` per GPU.
####################################################
import os

if "K_VALUES" in config:
    k_values = [str(x) for x in str(config["K_VALUES"]).split(",")]

if exp_type == 6:
    from khoice_b200 import pipeline6 as khb_p6
    if not os.path.isdir("exp6_input"):
        khb_p6.prepare_inputs(".", database_root, curr_trial, num_datasets)          # reference :31-57
    khb_p6.write_parse_time_files(".", k_values, num_datasets)                      # reference :27-29, 60-111

rule run_merge_list_exp6:
    """rules build_kmc_database_on_*_exp6 .. run_merge_list_exp6 of the reference for one k and both read types, fused"""
    input:
        lambda w: [khb_p6.p_reads(rt, n) for rt in khb_p6.READ_TYPES for n in range(1, num_datasets + 1)]
    output:
        expand("exp6_accuracies/{read_type}/values/k_{{k}}_accuracy_values.csv", read_type=["illumina", "ont"]),
        expand("exp6_accuracies/{read_type}/confusion_matrix/k_{{k}}_confusion_matrix.txt", read_type=["illumina", "ont"]),
        expand("exp6_accuracies/{read_type}/confusion_matrix/k_{{k}}_confusion_matrix_with_unidentified.txt", read_type=["illumina", "ont"])
    run:
        khb_p6.run_fused(".", num_datasets, [wildcards.k], trial=curr_trial)

rule concatenate_accuracies_exp6:
    input:
        expand("exp6_accuracies/{read_type}/values/k_{k}_accuracy_values.csv", k=k_values, read_type={"illumina", "ont"})
    output:
        f"exp6_accuracies/trial_{curr_trial}_short_acc.csv",
        f"exp6_accuracies/trial_{curr_trial}_long_acc.csv"
    shell:
        """
        printf "k,pivotnum,TP,TN,FP,FN,TP-U,TN-U,FP-U,FN-U\n" > {output[0]}
        printf "k,pivotnum,TP,TN,FP,FN,TP-U,TN-U,FP-U,FN-U\n" > {output[1]}
        cat exp6_accuracies/illumina/values/k_*_accuracy_values.csv >> {output[0]}
        cat exp6_accuracies/ont/values/k_*_accuracy_values.csv >> {output[1]}
        """

rule generate_exp6_output:
    input:
        f"exp6_accuracies/trial_{curr_trial}_short_acc.csv",
        f"exp6_accuracies/trial_{curr_trial}_long_acc.csv"
