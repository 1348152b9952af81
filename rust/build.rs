fn main() {
    // Point FUZZYBLUE_B200_LIB_DIR at fuzzyblue_b200/csrc (where `make` leaves libfuzzyblue_b200.so).
    if let Ok(dir) = std::env::var("FUZZYBLUE_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=fuzzyblue_b200");
}
