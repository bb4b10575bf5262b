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
#   * KHB_MODE=fused: ALL TEN rule names and output patterns stay (every reference target resolves, resume semantics and
#     downstream rules are unchanged), but the k-mer arithmetic runs in one GPU process per (k, group) (`khb group`, rule
#     within_group_union) and one per k (`khb across`, rule across_group_union).  step_1/2/7 databases are header-only
#     placeholders, the step_3 table is set-only (histogram + distinct k-mers), step_6 is a real set database (the hand-off
#     to the across-group job), and the three transform rules run the reference's own shell strings.  Run with `--cores 1`
#     per GPU, or KHB_WORKER_SOCKET: every GPU job owns the device.
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
    # fused: all ten rule names and every output pattern of the reference stay (any reference target still resolves), but the
    # arithmetic of the per-genome rules and of the two unions runs in ONE GPU process per (k, group) / per k:
    #   build_kmc_database_on_genome, transform_genome_to_set  -> header-only placeholders (no GPU, no k-mers)
    #   within_group_union    -> `khb group`: reads the group's genomes itself, writes the step_3 table set-only (histogram + k-mers)
    #   within_group_union_histogram, build_group_kmer_set, across_group_union_histogram -> the reference's own shell strings
    #   across_group_union    -> `khb across` over the step_6 sets; the step_7 table is header-only (histogram)
    rule build_kmc_database_on_genome:
        input: "data/dataset_{num}/{genome}.fna.gz"
        output: S1 + ".kmc_pre", S1 + ".kmc_suf"
        shell: "{KHB_BIN}/khb stub --k {wildcards.k} step_1/k_{wildcards.k}/dataset_{wildcards.num}/{wildcards.genome}"

    rule transform_genome_to_set:
        input: S1 + ".kmc_pre", S1 + ".kmc_suf"
        output: S2 + ".kmc_pre", S2 + ".kmc_suf"
        shell: "{KHB_BIN}/khb stub --k {wildcards.k} step_2/k_{wildcards.k}/dataset_{wildcards.num}/{wildcards.genome}.transformed"

    rule within_group_union:
        input: lambda w: [f"step_2/k_{w.k}/dataset_{w.num}/{g}.transformed.kmc_pre" for g in genomes_of(w.num)]
        output: S3 + ".kmc_pre", S3 + ".kmc_suf"
        params: genomes=lambda w: [f"data/dataset_{w.num}/{g}.fna.gz" for g in genomes_of(w.num)]
        shell: "{KHB_BIN}/khb group --k {wildcards.k} --table step_3/k_{wildcards.k}/dataset_{wildcards.num}/dataset_{wildcards.num}.transformed.combined {params.genomes}"

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
        params: sets=lambda w: [f"step_6/k_{w.k}/dataset_{n}/dataset_{n}.transformed.combined.transformed" for n in range(1, num_datasets + 1)]
        shell: "{KHB_BIN}/khb across --k {wildcards.k} --table step_7/k_{wildcards.k}/all_datasets.transformed.combined.transformed.combined {params.sets}"

    rule across_group_union_histogram:
        input: S7 + ".kmc_pre", S7 + ".kmc_suf"
        output: S8
        shell: "PATH={KHB_BIN}:$PATH kmc_tools transform step_7/k_{wildcards.k}/all_datasets.transformed.combined.transformed.combined histogram {output}"

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
