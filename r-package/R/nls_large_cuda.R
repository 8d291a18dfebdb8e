## nls_large_cuda.R -- R front-end of the B200 path of gsl_nls_large().
##
## Same user interface as the reference's gsl_nls_large() (R/nls_large.R:124-445 formula method, :459-627
## function method): formula / function + start + algorithm + control + jac + fvv + trace + weights, the same
## validation messages, the same control vectors (.ctrl_int[7], .ctrl_dbl[8], R/nls_large.R:383-407) and the
## same returned object (class c("gsl_nls", "nls"), R/nls_large.R:416-443).  What differs is what crosses the
## .Call boundary: R closures cannot run on a GPU, so instead of (.fn, .jac, .fvv, env) the model goes over as
## text -- the right-hand side of the formula, the parameter names and the predictor columns -- and is
## differentiated and compiled for the device by the library (deriv()-style symbolic Jacobian / directional second
## derivative, or the finite-difference rules of src/fdjac.c and src/fdfvv.c).
##
## Nothing of size n is computed or copied in R after the fit: residuals and the gradient stay on the device
## behind an external pointer and are materialised only when a method asks for them.

.cuda_algorithms <- c("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst")

## control list -> the two vectors of the reference boundary (R/nls_large.R:361-407)
.cuda_pack_control <- function(control, algorithm, trace) {
  ctrl <- gsl_nls_control()
  ctrl <- ctrl[grep("^mstart", names(ctrl), invert = TRUE)]
  if (!is.null(control)) {
    control <- as.list(control)
    ctrl[names(control)] <- control
  }
  ctrl$scale <- match.arg(ctrl$scale, c("more", "levenberg", "marquardt"))
  ctrl$solver <- "cholesky"
  ctrl$fdtype <- match.arg(ctrl$fdtype, c("forward", "center"))
  for (nm in c("maxiter", "factor_up", "factor_down", "avmax", "h_df", "h_fvv", "xtol", "ftol", "gtol")) {
    v <- ctrl[[nm]]
    if (!is.numeric(v) || length(v) != 1L || !(v > 0) || (nm == "maxiter" && v < 1))
      stop(sprintf("control$%s must be a positive number", nm))
  }
  list(
    ctrl = ctrl,
    int = c(maxiter = as.integer(ctrl$maxiter), trace = as.integer(isTRUE(trace)),
            algorithm = match(algorithm, .cuda_algorithms) - 1L,
            scale = match(ctrl$scale, c("more", "levenberg", "marquardt")) - 1L,
            fdtype = match(ctrl$fdtype, c("forward", "center")) - 1L,
            jacclass = -2L, jacnz = 0L),
    dbl = as.double(unlist(ctrl[c("factor_up", "factor_down", "avmax", "h_df", "h_fvv", "xtol", "ftol", "gtol")]))
  )
}

## jac / fvv arguments -> GSLNLS_JAC_* / GSLNLS_FVV_* (include/gslnls_b200.h)
.cuda_modes <- function(jac, fvv, algorithm, fdtype, weights_mode) {
  if (is.function(jac) || is.function(fvv))
    stop("R functions for 'jac' / 'fvv' cannot run on the GPU: use jac = TRUE (symbolic), \"forward\" or \"center\"")
  if (is.null(jac) || identical(jac, FALSE))
    stop("analytic Jacobian function 'jac' is required, but none is available")
  jm <- if (isTRUE(jac)) 0L else match(match.arg(jac, c("forward", "center")), c("forward", "center"))
  fm <- 0L
  if (identical(algorithm, "lmaccel")) {
    if (is.null(fvv) || identical(fvv, FALSE))
      stop("analytic second derivative function 'fvv' is required, but none is available")
    fm <- if (isTRUE(fvv)) 1L else 2L   # "fd": src/fdfvv.c
  }
  c(jm, fm, match(weights_mode, c("consistent", "gsl")) - 1L)
}

## the object of R/nls.R:1231-1411 (nlsModel), with the n-sized pieces lazy
.cuda_nlsModel <- function(form, cFit, lhs, wts, pnames) {
  cache <- new.env(parent = emptyenv())
  ev <- function(grad) {
    key <- if (grad) "rg" else "r"
    if (is.null(cache[[key]]))
      cache[[key]] <- .Call(C_nls_large_cuda_eval, cFit$handle, cFit$par, grad)
    cache[[key]]
  }
  swts <- if (length(wts)) sqrt(wts) else 1
  resid <- function() -ev(FALSE)$resid                       # R/nls.R:1255: resid <- -cFit$resid
  Rmat <- function() chol(cFit$jtj)                          # replaces qr.R(qr(.swts * gradient)), :1295
  list(
    resid = resid,
    fitted = function() lhs - resid() / swts,
    formula = function() form,
    deviance = function() cFit$ssr,
    lhs = function() lhs,
    gradient = function() { g <- ev(TRUE)$grad; colnames(g) <- pnames; g },
    conv = function() cFit$ssrtol,
    incr = function() drop(backsolve(Rmat(), backsolve(Rmat(), crossprod(ev(TRUE)$grad, resid()), transpose = TRUE))),
    getPars = function() cFit$par,
    getAllPars = function() cFit$par,
    getEnv = function() environment(form),
    trace = function() invisible(NULL),
    Rmat = Rmat,
    predict = function(newdata = list(), qr = FALSE) stop("use predict() on the fitted object"),
    release = function() invisible(.Call(C_nls_large_cuda_free, cFit$handle))
  )
}

gsl_nls_large_cuda <- function(fn, ...) UseMethod("gsl_nls_large_cuda")

gsl_nls_large_cuda.formula <- function(fn, data = parent.frame(), start,
                                       algorithm = c("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"),
                                       control = gsl_nls_control(), jac = TRUE, fvv = NULL, trace = FALSE,
                                       weights = NULL, devices = 0L, weights_mode = c("consistent", "gsl"), ...) {
  formula <- as.formula(fn)
  algorithm <- match.arg(algorithm)
  weights_mode <- match.arg(weights_mode)
  if (!is.list(data) && !is.environment(data))
    stop("'data' must be a list or an environment")
  if (missing(start))
    stop("starting values need to be provided")              # no selfStart evaluation on the device
  if (length(formula) == 2L) {                                # one-sided formula: response 0 (R/nls_large.R:150-153)
    formula[[3L]] <- formula[[2L]]
    formula[[2L]] <- 0
  }
  start <- unlist(start)
  pnames <- names(start)
  env <- if (!is.null(environment(formula))) environment(formula) else parent.frame()
  vnames <- setdiff(all.vars(formula[[3L]]), pnames)
  if (!length(vnames) && !length(pnames))
    stop("no parameters to fit and/or no data variables present")
  cols <- lapply(vnames, function(v) {
    x <- tryCatch(eval(as.name(v), data, env), error = function(e) NULL)
    if (is.null(x))
      stop(gettextf("parameters without starting value in 'data': %s", v), domain = NA)
    as.double(x)
  })
  lhs <- as.double(eval(formula[[2L]], data, env))
  n <- if (length(cols)) length(cols[[1L]]) else length(lhs)
  if (length(lhs) == 1L && n > 1L) lhs <- rep_len(lhs, n)
  if (any(vapply(cols, length, 0L) != length(lhs)))
    stop("variable lengths differ")
  if (length(lhs) < length(start))
    stop("negative residual degrees of freedom, cannot fit a model with less observations than parameters")
  if (!is.null(weights)) {
    weights <- as.double(weights)
    if (length(weights) != length(lhs))
      stop("'weights' should be numeric equal in length to 'y'")
    if (any(weights <= 0 | is.na(weights)))
      stop("missing or non-positive weights not allowed")
  }
  pk <- .cuda_pack_control(if (missing(control)) NULL else control, algorithm, trace)
  modes <- .cuda_modes(jac, fvv, algorithm, pk$ctrl$fdtype, weights_mode)
  rhs <- paste(deparse(formula[[3L]], width.cutoff = 500L), collapse = " ")

  cFit <- .Call(C_nls_large_cuda, rhs, pnames, vnames, cols, lhs, as.double(start), weights,
                pk$int, pk$dbl, as.integer(modes), as.integer(devices))

  m <- .cuda_nlsModel(formula, cFit, lhs, weights, pnames)
  convInfo <- list(isConv = as.logical(!cFit$conv), finIter = cFit$niter, finTol = cFit$ssrtol, nEval = cFit$neval,
                   trsName = paste("multilarge", cFit$algorithm, sep = "/"), stopCode = cFit$conv,
                   stopMessage = cFit$status)
  out <- list(m = m, data = substitute(data), convInfo = convInfo, call = match.call())
  out$call$algorithm <- algorithm
  out$call$control <- nls.control()
  out$call$trace <- isTRUE(trace)
  if (isTRUE(trace)) {
    keep <- seq_len(cFit$niter + 1L)
    out$partrace <- cFit$partrace[keep, , drop = FALSE]
    out$devtrace <- cFit$ssrtrace[keep]
  }
  out$control <- pk$ctrl
  if (!is.null(weights)) out$weights <- weights
  class(out) <- c("gsl_nls", "nls")
  out
}

## function method (R/nls_large.R:459-627): fn(par, ...) must be ONE arithmetic expression in par[i] /
## par[["name"]] / par["name"] and the names of `...`; it is rewritten into a formula right-hand side.
gsl_nls_large_cuda.function <- function(fn, y, start,
                                        algorithm = c("lm", "lmaccel", "dogleg", "ddogleg", "subspace2D", "cgst"),
                                        control = gsl_nls_control(), jac = TRUE, fvv = NULL, trace = FALSE,
                                        weights = NULL, devices = 0L, weights_mode = c("consistent", "gsl"), ...) {
  if (!is.numeric(y))
    stop("'y' should be a numeric response vector")
  if (missing(start) || is.null(names(unlist(start))))
    stop("starting values need to be provided as a named vector")
  arg <- names(formals(fn))[1L]
  pnames <- names(unlist(start))
  b <- body(fn)
  while (is.call(b) && identical(b[[1L]], as.name("{")) && length(b) == 2L) b <- b[[2L]]
  if (is.call(b) && identical(b[[1L]], as.name("{")))
    stop("the function method of the GPU path needs a function whose body is a single expression of its first argument")
  walk <- function(e) {
    if (is.call(e) && (identical(e[[1L]], as.name("[")) || identical(e[[1L]], as.name("[["))) &&
        identical(e[[2L]], as.name(arg))) {
      k <- e[[3L]]
      return(as.name(if (is.character(k)) k else pnames[[as.integer(k)]]))
    }
    if (is.call(e)) e[-1L] <- lapply(as.list(e)[-1L], walk)
    e
  }
  form <- as.formula(call("~", quote(.y), walk(b)), env = environment(fn))
  ## `...` are the extra arguments of fn (R/nls_large.R:459): they are the predictor columns here
  gsl_nls_large_cuda.formula(form, data = c(list(.y = y), list(...)), start = start, algorithm = algorithm,
                             control = control, jac = jac, fvv = fvv, trace = trace, weights = weights,
                             devices = devices, weights_mode = weights_mode)
}
