{
  "targets": [{
    "target_name": "genstark_b200",
    "sources": ["genstark_b200_addon.cc"],
    "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
    "libraries": ["-L<(module_root_dir)/../../genstark_b200", "-lgenstark_b200", "-Wl,-rpath,<(module_root_dir)/../../genstark_b200"],
    "defines": ["NAPI_CPP_EXCEPTIONS"],
    "cflags_cc": ["-std=c++17", "-fexceptions"]
  }]
}
