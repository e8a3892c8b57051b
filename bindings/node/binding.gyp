{
  "targets": [{
    "target_name": "montgomery_b200",
    "sources": ["addon.c"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../montgomery_b200", "-lmontgomery_b200", "-Wl,-rpath,<(module_root_dir)/../../montgomery_b200"]
  }]
}
