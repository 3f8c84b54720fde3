"""Copies the reference's own test fixtures (DATA, not sources) into tests/golden/test_data/ and records their
sha256, so that the parity tests can run where /root/reference does not exist (the GPU box).

Run in the build container:  python tests/golden/make_golden.py
Source: /root/reference/test_data (Helkafen/find-tfbs test_data/, used by src/main.rs:548-568 and src/bed.rs:67-95).
"""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/test_data"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_data")
FILES = ["ACGT.thr", "expected_output_1.vcf.gz", "expected_output_2.vcf.gz", "genotypes.bcf", "genotypes.bcf.csi",
         "genotypes2.bcf", "genotypes2.bcf.csi", "pwm_definitions.txt", "reference_genome.fa", "reference_genome.fa.fai",
         "regions1.bed", "regions2.bed", "samples"]

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    sums = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        sums[f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
    json.dump(sums, open(os.path.join(DST, "SHA256.json"), "w"), indent=1, sort_keys=True)
    print("copied", len(FILES), "fixtures")
