// examples/sample.rs of the reference (Rosenbrock N=100), with x and the objective on the device.
use lbfgs_b200::{lbfgs, DeviceBuffer, Rosenbrock};

fn main() -> anyhow::Result<()> {
    const N: usize = 100;
    let mut x0 = [0.0f64; N];
    for i in (0..N).step_by(2) { x0[i] = -1.2; x0[i + 1] = 1.0; }
    let mut x = DeviceBuffer::from_host(0, &x0)?;
    let prb = lbfgs().minimize(&mut x, Rosenbrock::new(0)?, |prgr| {
        println!("Iteration {}, Evaluation {}: fx = {:-12.6} xnorm = {:-12.6}, gnorm = {:-12.6}, ls = {}, step = {}",
                 prgr.niter, prgr.neval, prgr.fx, prgr.xnorm, prgr.gnorm, prgr.ncall, prgr.step);
        false
    })?;
    x.to_host(&mut x0)?;
    println!("fx = {:-12.6}, x[0] = {:-12.6}, x[1] = {:-12.6}", prb.fx, x0[0], x0[1]);
    Ok(())
}
