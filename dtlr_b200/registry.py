"""Name -> builder table with the public surface the reference's entry scripts use (models/registry.py:12-58):
`MODULE_BUILD_FUNCS.get(args.modelname)(args)` (finetuning.py:123-131), the decorator `@MODULE_BUILD_FUNCS.registe_with_name(
module_name="dino")` (models/dino/dino.py:1049) and plain `register(fn)`.  Written for this package: one dict, one decorator
factory; same names, same error types."""
import types


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}                  # builder name -> function(args) -> (model, criterion, postprocessors)

    def __len__(self):
        return len(self.module_dict)

    def __contains__(self, key):
        return key in self.module_dict

    def __repr__(self):
        return "Registry(name=%s, items=%s)" % (self.name, sorted(self.module_dict))

    def get(self, key):
        """the builder registered under `key`, or None (the reference's callers test for None)"""
        return self.module_dict.get(key)

    def register(self, module_build_function, module_name=None, force=False):
        if not isinstance(module_build_function, types.FunctionType):
            raise TypeError("module_build_function must be a function, but got %s" % type(module_build_function))
        key = module_name or module_build_function.__name__
        if key in self.module_dict and not force:
            raise KeyError("%s is already registered in %s" % (key, self.name))
        self.module_dict[key] = module_build_function
        return module_build_function

    def registe_with_name(self, module_name=None, force=False):      # (sic) spelled as in the reference, whose decorators we mirror
        def decorator(fn):
            return self.register(fn, module_name=module_name, force=force)
        return decorator


MODULE_BUILD_FUNCS = Registry("model build functions")
