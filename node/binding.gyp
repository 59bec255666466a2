{
  # node-gyp recipe for the N-API addon (run `npx node-gyp rebuild` in node/ on a machine with Node.js >= 12 and the built
  # ../wasmsnark_b200/libwasmsnark_b200.so).  Not runnable in the build image (no Node.js); the same source is linked and
  # executed there against a mock N-API runtime (tests/test_node_addon_mock.py).
  "targets": [
    {
      "target_name": "wasmsnark_b200",
      "sources": ["wasmsnark_b200_addon.c"],
      "include_dirs": ["../include"],
      "defines": ["NAPI_VERSION=6"],
      "libraries": ["-L<(module_root_dir)/../wasmsnark_b200", "-lwasmsnark_b200", "-Wl,-rpath,<(module_root_dir)/../wasmsnark_b200"],
      "cflags": ["-std=c11", "-Wall", "-Wextra"]
    }
  ]
}
