"""Synthetic input shapes shared by the parity tests (arguments of tools/synth_paf.cpp)."""
# name -> (generator args, non_skip_linkable variants to run)
SMALL = {
    "c1_small": (["--preset", "c1", "--scale", 0.03], [False, True]),
    "c2_small": (["--preset", "c2", "--scale", 0.012], [False, True]),
    "ties": (["--contigs", 40, "--blocks", 30, "--sd", 10, "--p_dup", 0.15, "--p_trans", 0.15, "--p_inv", 0.15, "--seed", 11], [False, True]),
    "overlappy": (["--contigs", 30, "--blocks", 60, "--sd", 20, "--p_dup", 0.1, "--p_trans", 0.1, "--p_inv", 0.1, "--p_ovl", 0.6,
                   "--p_cont", 0.1, "--seed", 12, "--lmin", 2000, "--lmax", 20000], [False, True]),
    "tiny": (["--contigs", 60, "--blocks", 8, "--sd", 6, "--p_dup", 0.2, "--p_trans", 0.2, "--p_inv", 0.2, "--seed", 13, "--lmin", 500,
              "--lmax", 5000, "--gap_max", 200], [False, True]),
    "singletons": (["--contigs", 50, "--blocks", 2, "--sd", 2, "--seed", 14, "--lmin", 500, "--lmax", 5000], [False]),
    "dense200": (["--preset", "c4", "--n", 200], [False, True]),
    "dense400": (["--preset", "c4", "--n", 400], [False, True]),
    # contigs of several buckets of 256 blocks: the segment-parallel relax (articulation blocks, tie-break conditions, sweep)
    "segments": (["--contigs", 6, "--blocks", 1500, "--sd", 300, "--p_dup", 0.1, "--p_trans", 0.1, "--p_inv", 0.1, "--seed", 31], [False, True]),
    "segments_long": (["--contigs", 3, "--blocks", 5000, "--sd", 1500, "--p_trans", 0.01, "--p_inv", 0.01, "--seed", 32], [False, True]),
    # one chain-like contig of the bench workloads' shape, still small enough for the reference's 56 n^2 B tables (0.5 GB)
    "chain3000": (["--contigs", 1, "--blocks", 3000, "--sd", 0, "--p_trans", 0.01, "--p_inv", 0.01, "--gap_max", 3000, "--lmin", 1000,
                   "--lmax", 12000, "--seed", 33], [False]),
    "cancer_small": (["--preset", "c3", "--scale", 0.004], [False]),
}
