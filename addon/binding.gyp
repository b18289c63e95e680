{
  "targets": [{
    "target_name": "olap_gpu",
    "sources": ["olap_napi.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../olap_in_memory_b200", "-lolapgpu", "-Wl,-rpath,<(module_root_dir)/../olap_in_memory_b200"],
    "cflags_cc": ["-std=c++17"]
  }]
}
