//! Dumps dalek-produced goldens for the 13 fixture statements: sha256(R1CSProof::to_bytes()), sha256(.coms text) and the
//! first 65 proof bytes, in the format of tests/golden/fixtures.json.  Call sites mirrored: /root/reference/src/prove.rs:45-81
//! (prove), /root/reference/src/commitments.rs:28,40 (blindings, patched per PATCH.md).  Source only: no toolchain in the image.
use bulletproofs_gadgets::prove::prove;
use sha2::{Digest, Sha256};
use std::fs;

const STEMS: [&str; 13] = ["example", "bounds_check", "equality", "inequality", "less_than", "merkle_tree", "mimc_hash",
                           "set_membership", "or", "or2", "or3", "or4", "or5"];

fn main() {
    let dir = std::env::args().nth(1).expect("usage: dalek_golden <fixture dir>");
    println!("{{");
    for (k, stem) in STEMS.iter().enumerate() {
        let read = |ext: &str| fs::read_to_string(format!("{}/{}.{}", dir, stem, ext)).expect("fixture file");
        let (inst, wtns, gad) = (read("inst"), read("wtns"), read("gadgets"));
        let mut coms = String::new();
        let proof = prove(stem, inst, wtns, gad, &mut coms).expect("prove");       // deterministic after PATCH.md
        let hex = |b: &[u8]| b.iter().map(|x| format!("{:02x}", x)).collect::<String>();
        println!(" \"{}\": {{\"proof_len\": {}, \"proof_sha256\": \"{}\", \"coms_sha256\": \"{}\", \"proof_head\": \"{}\"}}{}",
                 stem, proof.len(), hex(&Sha256::digest(&proof)), hex(&Sha256::digest(coms.as_bytes())), hex(&proof[..65]),
                 if k + 1 < STEMS.len() { "," } else { "" });
    }
    println!("}}");
}
