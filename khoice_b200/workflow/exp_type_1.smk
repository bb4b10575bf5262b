####################################################
# exp_type_1.smk -- B200 drop-in for workflow/rules/exp_type_1.smk of vshiv18/khoice.
#
# Same rule names, same inputs, same outputs (step_1 .. step_9, final_results_type1/) as the reference file it
# replaces (include it instead of rules/exp_type_1.smk at workflow/Snakefile:52; rule names are global).
# Differences:
#   * k_values may come from the config (K_VALUES: "7,9,31"); the default is the reference's literal list
#     (workflow/Snakefile:36).
#   * KHB_MODE=rules (default "fused"): every rule runs the reference's own shell string; put khoice_b200/bin first
#     on PATH and `kmc` / `kmc_tools` resolve to the B200 shims.  That also works with the UNMODIFIED reference file.
#   * KHB_MODE=fused: the per-genome and per-group rules collapse into ONE process per (k, group) (`khb group`) and
#     the across-group rules into one per k (`khb across`).  Rule names and outputs stay, so targets, resume
#     semantics and downstream rules are unchanged; step_1/2/3/7 databases are header-only stubs, step_6 is real
#     (it is the hand-off between the two processes).  Run with `--cores 1` per GPU: every job owns the device.
####################################################
import os

KHB_MODE = config.get("KHB_MODE", "fused")
KHB_BIN = os.path.join(config.get("KHB_REPO", "."), "khoice_b200", "bin")
if "K_VALUES" in config:
    k_values = [str(x) for x in str(config["K_VALUES"]).split(",")]

def genomes_of(num):
    return sorted(f.split(".fna.gz")[0] for f in os.listdir(f"data/dataset_{num}") if f.endswith(".fna.gz"))

def get_num_of_dataset_members(dataset_num):
    return len(genomes_of(dataset_num))

if exp_type == 1:
    from khoice_b200 import pipeline as khb_pipeline, tables as khb_tables
    khb_pipeline.write_complex_ops(".", k_values, num_datasets)   # tmp/ + complex_ops/ (reference :26-84)

S1 = "step_1/k_{k}/dataset_{num}/{genome}"
S2 = "step_2/k_{k}/dataset_{num}/{genome}.transformed"
S3 = "step_3/k_{k}/dataset_{num}/dataset_{num}.transformed.combined"
S4 = "step_4/k_{k}/dataset_{num}/dataset_{num}_k{k}_hist.txt"
S6 = "step_6/k_{k}/dataset_{num}/dataset_{num}.transformed.combined.transformed"
S7 = "step_7/k_{k}/all_datasets.transformed.combined.transformed.combined"
S8 = "step_8/k_{k}/all_datasets_k{k}_hist.txt"

if KHB_MODE == "rules":
    rule build_kmc_database_on_genome:
        input: "data/dataset_{num}/{genome}.fna.gz"
        output: S1 + ".kmc_pre", S1 + ".kmc_suf"
        shell: "PATH={KHB_BIN}:$PATH kmc -fm -m64 -k{wildcards.k} -ci1 {input} step_1/k_{wildcards.k}/dataset_{wildcards.num}/{wildcards.genome} tmp/"

    rule transform_genome_to_set:
        input: S1 + ".kmc_pre", S1 + ".kmc_suf"
        output: S2 + ".kmc_pre", S2 + ".kmc_suf"
        shell: "PATH={KHB_BIN}:$PATH kmc_tools transform step_1/k_{wildcards.k}/dataset_{wildcards.num}/{wildcards.genome} set_counts 1 step_2/k_{wildcards.k}/dataset_{wildcards.num}/{wildcards.genome}.transformed"

    rule within_group_union:
        input: lambda w: [f"step_2/k_{w.k}/dataset_{w.num}/{g}.transformed.kmc_pre" for g in genomes_of(w.num)]
        output: S3 + ".kmc_pre", S3 + ".kmc_suf"
        shell: "PATH={KHB_BIN}:$PATH kmc_tools complex complex_ops/within_groups/k_{wildcards.k}/dataset_{wildcards.num}/within_dataset_{wildcards.num}.txt"

    rule within_group_union_histogram:
        input: S3 + ".kmc_pre", S3 + ".kmc_suf"
        output: S4
        shell: "PATH={KHB_BIN}:$PATH kmc_tools transform step_3/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined histogram {output}"

    rule build_group_kmer_set:
        input: S3 + ".kmc_pre", S3 + ".kmc_suf"
        output: S6 + ".kmc_pre", S6 + ".kmc_suf"
        shell: "PATH={KHB_BIN}:$PATH kmc_tools transform step_3/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined set_counts 1 step_6/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined.transformed"

    rule across_group_union:
        input: lambda w: [f"step_6/k_{w.k}/dataset_{n}/dataset_{n}.transformed.combined.transformed.kmc_pre" for n in range(1, num_datasets + 1)]
        output: S7 + ".kmc_pre", S7 + ".kmc_suf"
        shell: "PATH={KHB_BIN}:$PATH kmc_tools complex complex_ops/across_groups/k_{wildcards.k}/across_all_datasets.txt"

    rule across_group_union_histogram:
        input: S7 + ".kmc_pre", S7 + ".kmc_suf"
        output: S8
        shell: "PATH={KHB_BIN}:$PATH kmc_tools transform step_7/k_{wildcards.k}/all_datasets.transformed.combined.transformed.combined histogram {output}"
else:
    # fused: the group job writes every per-genome / per-group output of its (k, group) in one go
    rule within_group_union_histogram:
        input: lambda w: [f"data/dataset_{w.num}/{g}.fna.gz" for g in genomes_of(w.num)]
        output: S4, S3 + ".kmc_pre", S3 + ".kmc_suf", S6 + ".kmc_pre", S6 + ".kmc_suf"
        shell: "{KHB_BIN}/khb group --k {wildcards.k} --hist {output[0]} --table step_3/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined --set step_6/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined.transformed {input}"

    rule across_group_union_histogram:
        input: lambda w: [f"step_6/k_{w.k}/dataset_{n}/dataset_{n}.transformed.combined.transformed.kmc_pre" for n in range(1, num_datasets + 1)]
        output: S8, S7 + ".kmc_pre", S7 + ".kmc_suf"
        params: sets=lambda w: [f"step_6/k_{w.k}/dataset_{n}/dataset_{n}.transformed.combined.transformed" for n in range(1, num_datasets + 1)]
        shell: "{KHB_BIN}/khb across --k {wildcards.k} --hist {output[0]} --table step_7/k_{wildcards.k}/all_datasets.transformed.combined.transformed.combined {params.sets}"

rule within_group_union_analysis:
    input: expand("step_4/k_{k_len}/dataset_{num}/dataset_{num}_k{k_len}_hist.txt", k_len=k_values, num=list(range(1, num_datasets + 1)))
    output: "step_5/within_datasets_analysis.csv"
    run: khb_tables.within_group_union_analysis(list(input), output[0], num_datasets, get_num_of_dataset_members)

rule across_group_union_analysis:
    input: expand("step_8/k_{k_len}/all_datasets_k{k_len}_hist.txt", k_len=k_values)
    output: "step_9/across_datasets_analysis.csv"
    run: khb_tables.across_group_union_analysis(list(input), output[0], num_datasets)

rule copy_final_results_type1:
    input: "step_5/within_datasets_analysis.csv", "step_9/across_datasets_analysis.csv"
    output: "final_results_type1/within_datasets_analysis.csv", "final_results_type1/across_datasets_analysis.csv"
    run:
        shell("cp {input[0]} {output[0]}")
        shell("cp {input[1]} {output[1]}")
