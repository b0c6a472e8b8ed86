{
  "targets": [{
    "target_name": "gridllm_native",
    "sources": ["addon.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../gridllm_b200", "-lgridllm_native", "-Wl,-rpath,<(module_root_dir)/../../gridllm_b200"],
    "cflags_cc": ["-std=c++17", "-O2"]
  }]
}
